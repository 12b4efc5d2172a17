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

import ctypes

import torch

from . import _lib, ops
from .sinkhorn import damping, scaling_parameters, sinkhorn_cost_batched

# largest cloud served by the one-launch-per-iteration kernels; beyond, the tiled TMA kernels win
# (measured crossover: profiles/r02_small_n.md)
SMALL_MAX = 6000


def eligible(N, M, D):
    return max(N, M) <= SMALL_MAX and D <= ops.MAX_D


def _views(buf, B, N, M, debias):
    f_ba = buf[: B * N].view(B, N)
    g_ab = buf[B * N: B * (N + M)].view(B, M)
    if not debias:
        return f_ba, g_ab, None, None
    return f_ba, g_ab, buf[B * (N + M): B * (2 * N + M)].view(B, N), buf[B * (2 * N + M):].view(B, M)


def _descent(x, y, a, b, eps_list, rho, p, debias):
    """Initialisation + every eps-scaling step in ONE C call (n_eps + 1 launches enqueued back to back); weights are
    passed as they are, their logarithm (with the reference's -100000 floor) is taken inside the kernel."""
    B, N, D = x.shape
    M = y.shape[1]
    dev = x.device
    bufs = torch.empty(2, B * (2 * N + 2 * M), dtype=torch.float32, device=dev)
    eps_arr = (ctypes.c_double * len(eps_list))(*[float(e) for e in eps_list])
    which = ctypes.c_int32(0)
    L = _lib.lib()
    P = ops._ptr
    with torch.cuda.device(dev):
        rc = L.b200ot_sinkhorn_loop_small(P(x), P(y), P(a), P(b), 1, eps_arr, len(eps_list),
                                          -1.0 if rho is None else float(rho), 1 if debias else 0, P(bufs[0]),
                                          P(bufs[1]), ctypes.byref(which), B, N, M, D, int(p), ops._stream(dev))
    _lib.check(rc, "b200ot_sinkhorn_loop_small")
    ops.count_launches(len(eps_list) + 1)
    return _views(bufs[0] if which.value else bufs[1], B, N, M, debias)


class _FinalStep(torch.autograd.Function):
    """The last, non-averaged update (sinkhorn_divergence.py:612-623): potentials detached, gradient to x and y only."""

    @staticmethod
    def forward(ctx, x, y, a, b, f_ba, g_ab, f_aa, g_bb, eps, lam, p, debias):
        B, N, D = x.shape
        M = y.shape[1]
        dev = x.device
        new = torch.empty(B * (2 * N + 2 * M), dtype=torch.float32, device=dev)
        n_f_ba, n_g_ab, n_f_aa, n_g_bb = _views(new, B, N, M, debias)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        lse2 = torch.empty(B * (2 * N + 2 * M), device=dev) if need else None
        L = _lib.lib()
        P = ops._ptr
        with torch.cuda.device(dev):
            rc = L.b200ot_sinkhorn_iteration_small(P(x), P(y), P(a), P(b), P(f_ba), P(g_ab), P(f_aa), P(g_bb),
                                                   P(n_f_ba), P(n_g_ab), P(n_f_aa), P(n_g_bb), P(lse2), B, N, M, D,
                                                   int(p), float(eps), 0.0, float(lam), 1, ops._stream(dev))
        _lib.check(rc, "b200ot_sinkhorn_iteration_small")
        ops.count_launches(1)
        if need:
            ctx.save_for_backward(x, y, a, b, f_ba, g_ab, f_aa if debias else f_ba, g_bb if debias else g_ab, lse2)
            ctx.meta = (float(eps), float(lam), int(p), bool(debias))
        if debias:
            return n_f_ba, n_g_ab, n_f_aa, n_g_bb
        return n_f_ba, n_g_ab

    @staticmethod
    def backward(ctx, *gos):
        x, y, a, b, f_ba, g_ab, f_aa, g_bb, lse2 = ctx.saved_tensors
        eps, lam, p, debias = ctx.meta
        B, N, D = x.shape
        M = y.shape[1]
        gos = [None if g is None else g.contiguous() for g in gos] + [None] * (4 - len(gos))
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        L = _lib.lib()
        P = ops._ptr
        with torch.cuda.device(x.device):
            rc = L.b200ot_sinkhorn_final_bwd_small(P(x), P(y), P(a), P(b), P(f_ba), P(g_ab),
                                                   P(f_aa) if debias else None, P(g_bb) if debias else None, P(lse2),
                                                   P(gos[0]), P(gos[1]), P(gos[2]), P(gos[3]), P(gx), P(gy), B, N, M,
                                                   D, p, eps, lam, None, 1, ops._stream(x.device))
        _lib.check(rc, "b200ot_sinkhorn_final_bwd_small")
        ops.count_launches(1)
        return (gx, gy) + (None,) * 10


class _FinalValue(torch.autograd.Function):
    """Final update + sinkhorn_cost (sinkhorn_divergence.py:612-623, :165-255) as TWO launches forward and ONE backward:
    the (B,) loss values come from b200ot_sinkhorn_cost_small, which also leaves d value / d potential for the backward
    kernel — no elementwise torch kernels, no autograd graph between the potentials and the value."""

    @staticmethod
    def forward(ctx, x, y, a, b, f_ba, g_ab, f_aa, g_bb, eps, lam, p, debias, rho, eps_final):
        B, N, D = x.shape
        M = y.shape[1]
        dev = x.device
        tot = B * (2 * N + 2 * M)
        need_xy = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        # one allocation: new potentials | lse2 | d value / d potential | phi, psi
        work = torch.empty(3 * tot + B * (N + M), dtype=torch.float32, device=dev)
        new, lse2, gos, terms = work[:tot], work[tot: 2 * tot], work[2 * tot: 3 * tot], work[3 * tot:]
        value = torch.empty(B, dtype=torch.float32, device=dev)
        n_f_ba, n_g_ab, n_f_aa, n_g_bb = _views(new, B, N, M, debias)
        go_f_ba, go_g_ab, go_f_aa, go_g_bb = _views(gos, B, N, M, debias)
        ad, bd = a.detach(), b.detach()
        L = _lib.lib()
        P = ops._ptr
        with torch.cuda.device(dev):
            st = ops._stream(dev)
            rc = L.b200ot_sinkhorn_iteration_small(P(x), P(y), P(ad), P(bd), P(f_ba), P(g_ab), P(f_aa), P(g_bb),
                                                   P(n_f_ba), P(n_g_ab), P(n_f_aa), P(n_g_bb),
                                                   P(lse2) if need_xy else None, B, N, M, D, int(p), float(eps), 0.0,
                                                   float(lam), 1, st)
            _lib.check(rc, "b200ot_sinkhorn_iteration_small")
            rc = L.b200ot_sinkhorn_cost_small(P(ad), P(bd), P(n_f_ba), P(n_g_ab), P(n_f_aa), P(n_g_bb), B, N, M,
                                              -1.0 if rho is None else float(rho), float(eps_final), P(value),
                                              P(go_f_ba) if need_xy else None, P(go_g_ab) if need_xy else None,
                                              P(go_f_aa) if need_xy and debias else None,
                                              P(go_g_bb) if need_xy and debias else None,
                                              P(terms[: B * N]) if need_w else None,
                                              P(terms[B * N:]) if need_w else None, st)
            _lib.check(rc, "b200ot_sinkhorn_cost_small")
        ops.count_launches(2)
        if need_xy or need_w:
            ctx.save_for_backward(x, y, ad, bd, f_ba, g_ab, f_aa if debias else f_ba, g_bb if debias else g_ab, work)
            ctx.meta = (float(eps), float(lam), int(p), bool(debias), need_xy, need_w)
        return value

    @staticmethod
    def backward(ctx, gv):
        x, y, a, b, f_ba, g_ab, f_aa, g_bb, work = ctx.saved_tensors
        eps, lam, p, debias, need_xy, need_w = ctx.meta
        B, N, D = x.shape
        M = y.shape[1]
        tot = B * (2 * N + 2 * M)
        gv = gv.contiguous().float()
        gx = gy = ga = gb = None
        if need_xy:
            lse2, gos = work[tot: 2 * tot], work[2 * tot: 3 * tot]
            go_f_ba, go_g_ab, go_f_aa, go_g_bb = _views(gos, B, N, M, debias)
            gx, gy = torch.empty_like(x), torch.empty_like(y)
            L = _lib.lib()
            P = ops._ptr
            with torch.cuda.device(x.device):
                rc = L.b200ot_sinkhorn_final_bwd_small(P(x), P(y), P(a), P(b), P(f_ba), P(g_ab),
                                                       P(f_aa) if debias else None, P(g_bb) if debias else None,
                                                       P(lse2), P(go_f_ba), P(go_g_ab), P(go_f_aa), P(go_g_bb), P(gx),
                                                       P(gy), B, N, M, D, p, eps, lam, P(gv), 1, ops._stream(x.device))
            _lib.check(rc, "b200ot_sinkhorn_final_bwd_small")
            ops.count_launches(1)
        if need_w:
            terms = work[3 * tot:]
            if ctx.needs_input_grad[2]:
                ga = terms[: B * N].view(B, N) * gv[:, None]
            if ctx.needs_input_grad[3]:
                gb = terms[B * N:].view(B, M) * gv[:, None]
        return (gx, gy, ga, gb) + (None,) * 10


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
    ad, bd = a.detach(), b.detach()
    with torch.no_grad():
        f_ba, g_ab, f_aa, g_bb = _descent(x.detach(), y.detach(), ad, bd, eps_list, rho, pk, debias)
    eps = eps_list[-1]
    lam = damping(eps, rho)
    if not potentials:
        return _FinalValue.apply(x, y, a, b, f_ba, g_ab, f_aa, g_bb, eps, lam, pk, debias, rho, eps_final)
    outs = _FinalStep.apply(x, y, ad, bd, f_ba, g_ab, f_aa, g_bb, eps, lam, pk, debias)
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
