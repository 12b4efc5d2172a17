"""Host side of the Sinkhorn divergence on point clouds: epsilon-scaling schedule, the symmetric
Sinkhorn loop and the dual-to-value formulas.  Every softmin lands in libb200ot.so (see ops.py).

This follows the control flow of the reference because the control flow *is* the specification
(src/geomloss/_legacy/sinkhorn_divergence.py:56-163 schedule/scalars, :171-250 value, :258-628 loop;
driver src/geomloss/_legacy/sinkhorn_samples.py:349-424), with three B200-side changes:
  * the "cost matrix" is the pair of point clouds — nothing of size N x M is ever stored;
  * ``h = log_w + pot/eps``, the damping factor and the ``1/2 (f + f~)`` averaging are fused into the
    softmin launch (prologue / epilogue) instead of separate elementwise kernels;
  * the autograd switch is a context manager, not the reference's global ``set_grad_enabled`` toggle.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def damping(eps, rho):
    """1 for balanced OT, 1/(1+eps/rho) with a KL marginal penalty.          sinkhorn_divergence.py:56-58"""
    return 1.0 if rho is None else 1.0 / (1.0 + eps / rho)


def log_weights(a):
    """log(a) with the log of non-positive weights pinned to -1e5.           sinkhorn_divergence.py:61-65"""
    return torch.where(a > 0, a.clamp_min(1e-45).log(), torch.full_like(a, -100000.0))


def max_diameter(x, y):
    """Length of the diagonal of the joint bounding box (one host sync).     sinkhorn_divergence.py:96-112"""
    lh = ops.cloud_extent(x, y)
    if lh is not None:  # one launch + one 2D-float copy; the norm of D numbers is taken on the host
        lo, hi = lh.cpu().double().unbind(0)
        return float(torch.sqrt(((hi - lo) ** 2).sum()).float())
    lo = torch.minimum(x.min(0).values, y.min(0).values)
    hi = torch.maximum(x.max(0).values, y.max(0).values)
    return (hi - lo).norm().item()


def epsilon_schedule(p, diameter, blur, scaling):
    """Temperatures diam^p -> blur^p, geometric with ratio scaling^p; the first value appears twice
    (the arange starts at p log diam) exactly as in the reference.          sinkhorn_divergence.py:115-151"""
    steps = np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))
    return [diameter**p] + [float(np.exp(e)) for e in steps] + [blur**p]


def scaling_parameters(x, y, p, blur, reach, diameter, scaling):
    """(diameter, eps, eps_list, rho).                                      sinkhorn_divergence.py:154-163"""
    if diameter is None:
        d = x.shape[-1]
        diameter = max_diameter(x.reshape(-1, d), y.reshape(-1, d))
    rho = None if reach is None else reach**p
    return diameter, blur**p, epsilon_schedule(p, diameter, blur, scaling), rho


def softmin_many(sm, calls):
    """Run the independent softmins ``calls = [(args, kwargs), ...]`` of one Jacobi iteration through ``sm``; an engine
    that can overlap their collectives (distributed.ColumnShardedEngine.softmin_raw_many) gets them all at once."""
    owner = getattr(sm, "__self__", None)
    many = getattr(owner, "softmin_raw_many", None) if owner is not None else None
    if many is not None:
        return many(calls)
    return [sm(*args, **kw) for args, kw in calls]


def sinkhorn_loop_points(a_log, b_log, x, y, eps_list, rho, *, p=2, debias=True, center=None, softmin_raw=None,
                         softmin_grad=None, problems=None):
    """Symmetric Sinkhorn iterations with eps-scaling on one pair of clouds.   sinkhorn_divergence.py:258-628

    The iterations run without autograd on detached clouds.  All four updates of an iteration read the
    OLD potentials (Jacobi), then  f <- 1/2 f + 1/2 lam softmin(...)  — one fused launch each.
    The last update is not averaged, takes ``x`` / ``y`` with autograd and detached right-hand sides,
    and therefore carries the whole gradient (envelope theorem).  Returns (f_aa, g_bb, g_ab, f_ba).

    ``softmin_raw`` / ``softmin_grad`` default to the single-GPU kernels; the column-sharded
    multi-GPU engine (distributed.py) injects its own pair with the same signatures.
    ``problems`` (batched inputs): ranges-mode descriptors {"xy", "yx", "xx", "yy"} of the block-diagonal problems
    obtained by stacking the B batch elements along the point axis — the whole batch then runs in ONE launch group
    per softmin, like the reference's batched LazyTensor reduction (sinkhorn_samples.py:229-290).
    """
    if problems is None:
        sm = softmin_raw or ops.softmin_raw
        smg = softmin_grad or ops.softmin
    else:
        from . import ranges

        def _prob(rows, cols):
            key = ("x" if rows is x or rows is xd else "y") + ("x" if cols is x or cols is xd else "y")
            # (x is y: every key reads "xx", and the four problems have the same shape)
            return problems[key] if key in problems else problems["xy"]

        def sm(eps, rows, cols, h_a, h_b=None, h_scale_b=0.0, **kw):
            return ranges.softmin_ranges_raw(eps, rows, cols, h_a, h_b, h_scale_b, _prob(rows, cols), **kw)

        def smg(eps, rows, cols, h_a, h_b=None, h_scale_b=0.0, **kw):
            return ranges.softmin_ranges(eps, rows, cols, h_a, h_b, h_scale_b, _prob(rows, cols), **kw)

    xd, yd = x.detach(), y.detach()
    with torch.no_grad():
        eps = eps_list[0]
        lam = damping(eps, rho)
        g_ab = sm(eps, yd, xd, a_log, p=p, center=center, beta=lam)[0]
        f_ba = sm(eps, xd, yd, b_log, p=p, center=center, beta=lam)[0]
        if debias:
            f_aa = sm(eps, xd, xd, a_log, p=p, center=center, beta=lam)[0]
            g_bb = sm(eps, yd, yd, b_log, p=p, center=center, beta=lam)[0]
        for eps in eps_list:
            lam = damping(eps, rho)
            inv = 1.0 / eps
            ukw = dict(p=p, center=center, alpha_old=0.5, beta=0.5 * lam)
            calls = [((eps, xd, yd, b_log, g_ab, inv), dict(out_old=f_ba, **ukw)),
                     ((eps, yd, xd, a_log, f_ba, inv), dict(out_old=g_ab, **ukw))]
            if debias:
                calls += [((eps, xd, xd, a_log, f_aa, inv), dict(out_old=f_aa, **ukw)),
                          ((eps, yd, yd, b_log, g_bb, inv), dict(out_old=g_bb, **ukw))]
            res = softmin_many(sm, calls)  # all four read the OLD potentials (Jacobi): independent
            f_ba, g_ab = res[0][0], res[1][0]
            if debias:
                f_aa, g_bb = res[2][0], res[3][0]
    # final extrapolation: eps / lam are the last values of the schedule (reference: leaked loop variables)
    inv = 1.0 / eps
    new_f_ba = smg(eps, x, yd, b_log, g_ab, inv, p=p, center=center, scale_out=lam)
    new_g_ab = smg(eps, y, xd, a_log, f_ba, inv, p=p, center=center, scale_out=lam)
    if debias:
        new_f_aa = smg(eps, x, xd, a_log, f_aa, inv, p=p, center=center, scale_out=lam)
        new_g_bb = smg(eps, y, yd, b_log, g_bb, inv, p=p, center=center, scale_out=lam)
        return new_f_aa, new_g_bb, new_g_ab, new_f_ba
    return None, None, new_g_ab, new_f_ba


def _dot(w, f):
    """<w, f> as an elementwise product + sum (a plain reduction kernel: no cuBLAS on the value path)."""
    return (w * f).sum()


def sinkhorn_cost(eps, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias=True, potentials=False):
    """Value of the divergence (or the dual potentials) from the four potentials, unbatched vectors.
    sinkhorn_divergence.py:171-250.  The unbalanced weight is (rho + eps/2) in both passes — the
    reference's UnbalancedWeight.backward is dead code."""
    if potentials:
        return (f_ba - f_aa, g_ab - g_bb) if debias else (f_ba, g_ab)
    if rho is None:
        if debias:
            return _dot(a, f_ba - f_aa) + _dot(b, g_ab - g_bb)
        return _dot(a, f_ba) + _dot(b, g_ab)
    w = rho + eps / 2
    if debias:
        return _dot(a, w * ((-f_aa / rho).exp() - (-f_ba / rho).exp())) + _dot(
            b, w * ((-g_bb / rho).exp() - (-g_ab / rho).exp()))
    return _dot(a, w * (1 - (-f_ba / rho).exp())) + _dot(b, w * (1 - (-g_ab / rho).exp()))


def sinkhorn_points(a, x, b, y, p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, debias=True,
                    potentials=False, softmin_raw=None, softmin_grad=None, keops=False, **_ignored):
    """Sinkhorn divergence between two weighted clouds a:(N,) x:(N,D) b:(M,) y:(M,D) on one CUDA device.
    Counterpart of sinkhorn_tensorized / sinkhorn_online (sinkhorn_samples.py:74-221, :349-424) for a single batch
    element.  ``keops``: the "online" cost convention — for p = 1, Norm2(x-y) without the 1e-8 clamp of the
    tensorized `distances` (sinkhorn_samples.py:303-306 vs utils.py:56-61)."""
    if p not in (1, 2):
        raise KeyError(p)  # the reference's cost table only knows p = 1, 2 (sinkhorn_samples.py:26-29)
    diameter, eps, eps_list, rho = scaling_parameters(x, y, p, blur, reach, diameter, scaling)
    center = ops.default_center(x.detach(), y.detach())
    pk = (p | ops.P_UNCLAMPED) if keops else p
    f_aa, g_bb, g_ab, f_ba = sinkhorn_loop_points(log_weights(a.detach()), log_weights(b.detach()), x, y, eps_list,
                                                  rho, p=pk, debias=debias, center=center, softmin_raw=softmin_raw,
                                                  softmin_grad=softmin_grad)
    return sinkhorn_cost(eps, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias=debias, potentials=potentials)


def _bsum(w, f, B):
    return (w * f).view(B, -1).sum(1)


def sinkhorn_cost_batched(eps, rho, a, b, f_aa, g_bb, g_ab, f_ba, B, debias=True, potentials=False):
    """sinkhorn_cost on B problems stacked along the point axis: one value per batch element (scal(batch=True),
    utils.py:13-18)."""
    if potentials:
        return (f_ba - f_aa, g_ab - g_bb) if debias else (f_ba, g_ab)
    if rho is None:
        if debias:
            return _bsum(a, f_ba - f_aa, B) + _bsum(b, g_ab - g_bb, B)
        return _bsum(a, f_ba, B) + _bsum(b, g_ab, B)
    w = rho + eps / 2
    if debias:
        return _bsum(a, w * ((-f_aa / rho).exp() - (-f_ba / rho).exp()), B) + _bsum(
            b, w * ((-g_bb / rho).exp() - (-g_ab / rho).exp()), B)
    return _bsum(a, w * (1 - (-f_ba / rho).exp()), B) + _bsum(b, w * (1 - (-g_ab / rho).exp()), B)


def sinkhorn_points_batched(a, x, b, y, p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, debias=True,
                            potentials=False, keops=False, **_ignored):
    """Batched Sinkhorn divergence a:(B,N) x:(B,N,D) b:(B,M) y:(B,M,D), D <= 8: every softmin of the loop is ONE
    ranges-mode launch group over the whole batch (block-diagonal problem), not B launches.
    Reference: sinkhorn_tensorized / sinkhorn_online on (B, ...) inputs, softmin_online_lazytensor
    (sinkhorn_samples.py:74-221, :229-290, :349-424); one shared eps-schedule (diameter over the flattened batch)."""
    from . import ranges

    if p not in (1, 2):
        raise KeyError(p)
    B, N, D = x.shape
    M = y.shape[1]
    diameter, eps, eps_list, rho = scaling_parameters(x, y, p, blur, reach, diameter, scaling)
    xf, yf, af, bf = x.reshape(B * N, D), y.reshape(B * M, D), a.reshape(B * N), b.reshape(B * M)
    center = ops.default_center(xf.detach(), yf.detach())
    dev = x.device
    problems = {"xy": ranges.batch_problem(B, N, M, dev), "yx": ranges.batch_problem(B, M, N, dev)}
    if debias:
        problems["xx"] = ranges.batch_problem(B, N, N, dev)
        problems["yy"] = ranges.batch_problem(B, M, M, dev)
    pk = (p | ops.P_UNCLAMPED) if keops else p
    f_aa, g_bb, g_ab, f_ba = sinkhorn_loop_points(log_weights(af.detach()), log_weights(bf.detach()), xf, yf,
                                                  eps_list, rho, p=pk, debias=debias, center=center,
                                                  problems=problems)
    out = sinkhorn_cost_batched(eps, rho, af, bf, f_aa, g_bb, g_ab, f_ba, B, debias=debias, potentials=potentials)
    if potentials:
        return out[0].view(B, N), out[1].view(B, M)
    return out
