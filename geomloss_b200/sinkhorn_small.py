"""Sinkhorn divergence on SMALL point clouds (N, M up to a few thousand, batched or not): one kernel launch per
symmetric Sinkhorn iteration instead of twelve.

Same algorithm and autograd contract as sinkhorn.sinkhorn_points / the reference's sinkhorn_tensorized and
sinkhorn_online (src/geomloss/_legacy/sinkhorn_samples.py:74-221, :349-424; loop sinkhorn_divergence.py:258-628),
driven through ``b200ot_sinkhorn_iteration_small`` (csrc/b200ot_small.cu): the four softmins of an iteration — which
all read the OLD potentials — are the y-slices of one grid, batch elements its z-slices; initialisation, every
eps-scaling step and the final gradient-carrying step are one launch each, the gradient w.r.t. both clouds one more.
The reference's own benchmark protocol at N = 1 000 (blur = .05: 8 temperatures) is then 10 + 1 launches.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .sinkhorn import damping, log_weights, scaling_parameters, sinkhorn_cost_batched

# largest cloud served by the one-launch-per-iteration kernels; beyond, the tiled TMA kernels win
# (measured crossover: profiles/r02_small_n.md)
SMALL_MAX = 6000


def eligible(N, M, D):
    return max(N, M) <= SMALL_MAX and D <= ops.MAX_D


def _iteration(x, y, a_log, b_log, pots, outs, eps, alpha_old, beta, p, lse2=None):
    """pots: (f_ba, g_ab, f_aa, g_bb) or None (initialisation); outs: 4 tensors (f_aa / g_bb None without debias)."""
    B, N, D = x.shape
    M = y.shape[1]
    L = _lib.lib()
    P = ops._ptr
    f_ba, g_ab, f_aa, g_bb = pots if pots is not None else (None, None, None, None)
    with torch.cuda.device(x.device):
        rc = L.b200ot_sinkhorn_iteration_small(P(x), P(y), P(a_log), P(b_log), P(f_ba), P(g_ab), P(f_aa), P(g_bb),
                                               P(outs[0]), P(outs[1]), P(outs[2]), P(outs[3]), P(lse2), B, N, M, D,
                                               int(p), float(eps), float(alpha_old), float(beta),
                                               ops._stream(x.device))
    _lib.check(rc, "b200ot_sinkhorn_iteration_small")
    ops.count_launches(1)


class _FinalStep(torch.autograd.Function):
    """The last, non-averaged update (sinkhorn_divergence.py:612-623): potentials detached, gradient to x and y only."""

    @staticmethod
    def forward(ctx, x, y, a_log, b_log, f_ba, g_ab, f_aa, g_bb, eps, lam, p, debias):
        B, N, D = x.shape
        M = y.shape[1]
        dev = x.device
        new = [torch.empty(B, N, device=dev), torch.empty(B, M, device=dev),
               torch.empty(B, N, device=dev) if debias else None, torch.empty(B, M, device=dev) if debias else None]
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        lse2 = torch.empty(B * (2 * N + 2 * M), device=dev) if need else None
        _iteration(x, y, a_log, b_log, (f_ba, g_ab, f_aa, g_bb), new, eps, 0.0, lam, p, lse2=lse2)
        if need:
            ctx.save_for_backward(x, y, a_log, b_log, f_ba, g_ab, f_aa if debias else f_ba, g_bb if debias else g_ab, lse2)
            ctx.meta = (float(eps), float(lam), int(p), bool(debias))
        if debias:
            return new[0], new[1], new[2], new[3]
        return new[0], new[1]

    @staticmethod
    def backward(ctx, *gos):
        x, y, a_log, b_log, f_ba, g_ab, f_aa, g_bb, lse2 = ctx.saved_tensors
        eps, lam, p, debias = ctx.meta
        B, N, D = x.shape
        M = y.shape[1]
        gos = [None if g is None else g.contiguous() for g in gos] + [None] * (4 - len(gos))
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        L = _lib.lib()
        P = ops._ptr
        with torch.cuda.device(x.device):
            rc = L.b200ot_sinkhorn_final_bwd_small(P(x), P(y), P(a_log), P(b_log), P(f_ba), P(g_ab),
                                                   P(f_aa) if debias else None, P(g_bb) if debias else None, P(lse2),
                                                   P(gos[0]), P(gos[1]), P(gos[2]), P(gos[3]), P(gx), P(gy), B, N, M,
                                                   D, p, eps, lam, ops._stream(x.device))
        _lib.check(rc, "b200ot_sinkhorn_final_bwd_small")
        ops.count_launches(1)
        return (gx, gy) + (None,) * 10


def sinkhorn_small(a, x, b, y, p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, debias=True, potentials=False,
                   keops=False, **_ignored):
    """a:(B,N) x:(B,N,D) b:(B,M) y:(B,M,D) float32 CUDA tensors -> (B,) values or the (B,N), (B,M) potentials."""
    if p not in (1, 2):
        raise KeyError(p)
    x, y = ops._f32c(x, "x"), ops._f32c(y, "y")
    a, b = ops._f32c(a, "a"), ops._f32c(b, "b")
    B, N, D = x.shape
    M = y.shape[1]
    diameter, eps_final, eps_list, rho = scaling_parameters(x, y, p, blur, reach, diameter, scaling)
    pk = (p | ops.P_UNCLAMPED) if keops else p
    a_log, b_log = log_weights(a.detach()), log_weights(b.detach())
    xd, yd = x.detach(), y.detach()
    dev = x.device

    def fresh():
        return [torch.empty(B, N, device=dev), torch.empty(B, M, device=dev),
                torch.empty(B, N, device=dev) if debias else None, torch.empty(B, M, device=dev) if debias else None]

    cur, nxt = fresh(), fresh()
    with torch.no_grad():
        eps = eps_list[0]
        _iteration(xd, yd, a_log, b_log, None, cur, eps, 0.0, damping(eps, rho), pk)
        for eps in eps_list:
            lam = damping(eps, rho)
            _iteration(xd, yd, a_log, b_log, cur, nxt, eps, 0.5, 0.5 * lam, pk)
            cur, nxt = nxt, cur
    f_ba, g_ab, f_aa, g_bb = cur
    outs = _FinalStep.apply(x, y, a_log, b_log, f_ba, g_ab, f_aa, g_bb, eps, lam, pk, debias)
    if debias:
        n_f_ba, n_g_ab, n_f_aa, n_g_bb = outs
    else:
        (n_f_ba, n_g_ab), n_f_aa, n_g_bb = outs, None, None
    out = sinkhorn_cost_batched(eps_final, rho, a.reshape(-1), b.reshape(-1),
                                None if n_f_aa is None else n_f_aa.reshape(-1),
                                None if n_g_bb is None else n_g_bb.reshape(-1), n_g_ab.reshape(-1),
                                n_f_ba.reshape(-1), B, debias=debias, potentials=potentials)
    if potentials:
        return out[0].view(B, N), out[1].view(B, M)
    return out
