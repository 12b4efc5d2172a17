"""``geomloss.ot.solve_sample`` over the B200 softmin kernels.

The reference's "new API" for point clouds (src/geomloss/ot/_implementations/sample.py:190-395) builds dense (N, M)
cost matrices and runs ``sinkhorn_loop`` (ot/_abstract_solvers/sinkhorn_ot.py:240-447) on them.  Here the same
iteration runs on the fused on-the-fly softmin of ``ops.softmin_raw`` (never an N x M matrix), so the facade works
at N = M = 1e6 like the rest of the engine.  Conventions of THIS API (they differ from ``SamplesLoss``):

* cost ``"sqeuclidean"`` is ``C(x, y) = |x - y|^2`` — no 1/2 (sample.py:38-66); ``reg`` is the temperature eps and
  ``unbalanced`` the marginal penalty rho, used as given; ``blur`` / ``reach`` are shortcuts for ``2 blur^2`` /
  ``2 reach^2`` (sample.py:283-297);
* the eps ladder is ``geomspace(diameter^2, reg, max_iter)`` (annealing.py:131-170), one symmetric averaged update
  per value and one final non-averaged update (sinkhorn_ot.py:262-288, :419-436);
* the first iterate is the eps = +inf softmin (cost averages), minus the per-point ``0.5 a_i f_i`` — the
  reference's ``bk.dot_products`` treats the N axis of un-batched vectors as a batch axis (sinkhorn_ot.py:17-29,
  _backends/torch.py:28-32); kept, the goldens pin it.

Mapping onto the kernel (which evaluates ``-e log sum_j exp(h_j - (|x-y|^2/2)/e)``): with ``e = reg/2``,
``-reg log sum_j b_j exp((g_j - |x-y|^2)/reg) = 2 * softmin_kernel(e; h = log b + g/reg)``.

Autograd: like the reference, ``result.value`` is differentiable w.r.t. ``X_a``, ``X_b`` (through the cost in the
LAST update only — log-weights and incoming potentials are detached, sinkhorn_ot.py:419-436) and w.r.t. ``a``, ``b``
(their direct appearance in the value formula).  Unlike ``SamplesLoss`` the last update is differentiated w.r.t.
BOTH clouds of every softmin: the column gradient  sum_i u_i p_ij 2 (y_j - x_i)  is the row-gradient kernel run on
the swapped problem with column "log-weights" log u_i + softmin_i/eps (see ``_LastUpdate``).  The other attributes
(potentials, marginals, plan) are returned detached.  Inputs: float32 CUDA tensors; no CPU path.
Parity: pinned — ``tests/golden/ot_sample_case*.npz`` hold fp32 and fp64 runs of the real reference.
"""
from __future__ import annotations

import math
from functools import cached_property

import numpy as np
import torch

from . import ops
from .sinkhorn import log_weights, max_diameter

__all__ = ["solve_sample", "solve_sample_batch", "OTResultSample", "LinearOperator"]

_DENSE_LIMIT = 1 << 27  # entries of a dense (N, M) plan we are willing to materialise on request


# ----------------------------------------------------------------------------------------------------
# argument checks                                            src/geomloss/_arguments.py:14-154
# ----------------------------------------------------------------------------------------------------
def _check_regularization(reg, unbalanced, unbalanced_type, method, tol, max_iter):
    if reg is None:
        raise TypeError("'<' not supported between instances of 'NoneType' and 'int'")  # the reference's own failure
    if reg < 0:
        raise ValueError(f"Parameter 'reg' should be >= 0. Received {reg}.")
    if reg == 0:
        raise NotImplementedError("Currently, we require that reg > 0.")
    if unbalanced is not None and unbalanced <= 0:
        raise ValueError(f"Parameter 'unbalanced' should be None (= +infty) or > 0. Received {unbalanced}.")
    if unbalanced_type != "KL":
        raise NotImplementedError("Currently, we only support unbalanced OT with a 'KL' penalty on the marginal "
                                  "constraints.")
    if method != "auto":
        raise NotImplementedError("Currently, we only support a single method.")
    if max_iter is None:
        raise ValueError("The 'max_iter' parameter should be a positive integer.")
    if tol is not None:
        raise NotImplementedError("Currently, we do not support rigorous stopping criteria.")


def _check_marginal(m, like, size, name):
    if m is None:
        return torch.full_like(like, 1.0 / size)
    if m.shape != like.shape:
        raise ValueError(f"The marginal '{name}' should be of shape {like.shape}. Instead, received an array of "
                         f"shape {m.shape}.")
    if bool((m < 0).any()):
        raise ValueError(f"The marginal '{name}' contains negative values. We require that {name} >= 0.")
    return m


def annealing_eps(maxmin_cost, eps, n_iter):
    """The eps ladder of ``annealing_parameters(maxmin_cost=, eps=, n_iter=)``, scaling=None (annealing.py:114-170)."""
    if n_iter <= 0:
        raise ValueError(f"The number of iterations should be >= 1. Received n_iter={n_iter}.")
    maxmin_cost = max(float(maxmin_cost), eps)
    if n_iter == 1:
        return [eps]
    return [float(e) for e in np.geomspace(maxmin_cost, eps, n_iter)]


# ----------------------------------------------------------------------------------------------------
# the solver
# ----------------------------------------------------------------------------------------------------
class _Softmin:
    """``softmin_sample`` (sample.py:91-182) for the four (rows, columns) pairs of one problem, on the kernel."""

    def __init__(self, X_a, X_b):
        self.center = ops.default_center(X_a, X_b)
        self.X_a, self.X_b = X_a, X_b

    def __call__(self, eps, x, y, log_w, pot, *, damp=1.0, old=None):
        """damp * softmin(eps, log_w, C(x, y), pot), or its average with ``old`` (the symmetric Sinkhorn update)."""
        if old is None:
            return ops.softmin_raw(0.5 * eps, x, y, log_w, pot, 1.0 / eps, p=2, center=self.center, beta=2.0 * damp)[0]
        return ops.softmin_raw(0.5 * eps, x, y, log_w, pot, 1.0 / eps, p=2, center=self.center, out_old=old,
                               alpha_old=0.5, beta=damp)[0]

    def at_infinity(self, x, y, w_rows, w_cols, damp):
        """eps = +inf: f_i = sum_j w_j |x_i - y_j|^2 / sum_j w_j by moments (fp64, centred), then the reference's
        offset and dampening (sinkhorn_ot.py:17-29)."""
        c = self.center.double()
        xc, yc, w = x.double() - c, y.double() - c, w_cols.double()
        W = w.sum()
        ybar = (w[:, None] * yc).sum(0) / W
        f = (xc * xc).sum(1) - 2.0 * (xc @ ybar) + (w * (yc * yc).sum(1)).sum() / W
        return (damp * (f - 0.5 * w_rows.double() * f)).float()


class _LastUpdate(torch.autograd.Function):
    """``damp * softmin(eps, log_w, C(rows, cols), pot)`` of the new API, differentiable w.r.t. rows AND cols.

    With p_ij the softmax weights of row i over the columns and u = damp * grad_output:
      d/d rows_i = u_i 2 (rows_i - sum_j p_ij cols_j)                    -> b200ot_softmin_bwd_x as is;
      d/d cols_j = sum_i u_i p_ij 2 (cols_j - rows_i) = 2 W_j (cols_j - sum_i q_ji rows_i),
         W_j = sum_i u_i p_ij = exp(log_w_j + pot_j/eps) sum_i exp(log u_i + softmin_i/eps - C_ij/eps),
      i.e. the SAME row-gradient kernel on the swapped problem (rows <- cols, column log-weights log u_i, column
      potential softmin_i), one extra softmin for its normaliser W; signed u is split into its two parts."""

    @staticmethod
    def forward(ctx, rows, cols, log_w, pot, eps, damp, center):
        out, lse2 = ops.softmin_raw(0.5 * eps, rows.detach(), cols.detach(), log_w, pot, 1.0 / eps, p=2, center=center,
                                    beta=2.0 * damp, want_lse2=True)
        ctx.save_for_backward(rows.detach(), cols.detach(), log_w, pot, out, lse2, center)
        ctx.meta = (float(eps), float(damp))
        return out

    @staticmethod
    def backward(ctx, go):
        rows, cols, log_w, pot, out, lse2, center = ctx.saved_tensors
        eps, damp = ctx.meta
        e2, inv = 0.5 * eps, 1.0 / eps
        u_all = (damp * go).float().contiguous()
        g_rows = g_cols = None
        if ctx.needs_input_grad[0]:
            g_rows = ops.softmin_grad_rows(e2, rows, cols, log_w, pot, inv, lse2, 2.0 * u_all, p=2, center=center)
        if ctx.needs_input_grad[1]:
            sm_und = out / damp  # undamped softmin values, new-API scale
            g_cols = torch.zeros_like(cols)
            for sign in (1.0, -1.0):
                u = (sign * u_all).clamp_min(0.0)
                if not bool((u > 0).any()):
                    continue
                lu = log_weights(u)
                sm_t, lse2_t = ops.softmin_raw(e2, cols, rows, lu, sm_und, inv, p=2, center=center, want_lse2=True)
                W = torch.exp((log_w.double() + pot.double() * inv - sm_t.double() / e2)).float()
                g_cols = g_cols + sign * ops.softmin_grad_rows(e2, cols, rows, lu, sm_und, inv, lse2_t,
                                                               (2.0 * W).contiguous(), p=2, center=center)
        return g_rows, g_cols, None, None, None, None, None


def solve_sample(X_a, X_b, a=None, b=None, cost="sqeuclidean", debias=False, reg=None, unbalanced=None,
                 unbalanced_type="KL", method="auto", max_iter=None, tol=None, blur=None, reach=None):
    """Entropic (un)balanced OT between two point clouds; mirrors ``geomloss.ot.solve_sample`` (sample.py:190-395).

    X_a: (N, D), X_b: (M, D) float32 CUDA tensors; a: (N,), b: (M,) non-negative weights (uniform 1/N, 1/M if None).
    Returns an :class:`OTResultSample`.
    """
    p = 2 if cost == "sqeuclidean" else 1
    if blur is not None:
        if reg is not None:
            raise ValueError("Parameters 'reg' and 'blur' are redundant. Please specify only one of them.")
        reg = p * (blur**p)
    if reach is not None:
        if unbalanced is not None:
            raise ValueError("Parameters 'unbalanced' and 'reach' are redundant. Please specify only one of them.")
        unbalanced = p * (reach**p)
    _check_regularization(reg, unbalanced, unbalanced_type, method, tol, max_iter)
    if not (torch.is_tensor(X_a) and torch.is_tensor(X_b)):
        raise TypeError("X_a and X_b must be torch CUDA tensors (this engine has no NumPy / CPU path)")
    if X_a.dim() != 2:
        raise ValueError(f"Expected X_a to be a (N, D) array. Received {tuple(X_a.shape)}.")
    if X_b.dim() != 2:
        raise ValueError(f"Expected X_b to be a (M, D) array. Received {tuple(X_b.shape)}.")
    N, D = X_a.shape
    M, D_ = X_b.shape
    if D != D_:
        raise ValueError("Expected X_a and X_b to have the same number of coordinates per sample. "
                         f"Received D={D} for X_a and D={D_} for X_b.")
    a = _check_marginal(a, X_a[:, 0], N, "a")
    b = _check_marginal(b, X_b[:, 0], M, "b")
    if unbalanced is None:
        sa, sb = float(a.detach().sum()), float(b.detach().sum())
        if abs(sa - sb) / (sa + sb) > 1e-3:
            raise ValueError("The two arrays of marginal weights 'a' and 'b' do not sum up to the same value. As a "
                             "consequence, the balanced OT problem is not feasible. To fix this error, you may either "
                             "normalize the two marginals or use UNbalanced optimal transport with the 'unbalanced' "
                             "keyword argument.")
    if cost != "sqeuclidean":
        raise NotImplementedError()  # as the reference (sample.py:76-88)
    reg = float(reg)
    rho = None if unbalanced is None else float(unbalanced)

    live = (X_a, X_b, a, b)  # the caller's tensors: the value stays attached to them when they require grad
    want_grad = torch.is_grad_enabled() and any(t.requires_grad for t in live)
    with torch.no_grad():
        X_a, X_b, a, b = X_a.detach(), X_b.detach(), a.detach(), b.detach()
        eps_list = annealing_eps(max_diameter(X_a, X_b) ** p, reg, int(max_iter))
        sm = _Softmin(X_a, X_b)
        log_a, log_b = log_weights(a), log_weights(b)

        def damp(eps):
            return 1.0 if rho is None else 1.0 / (1.0 + eps / rho)

        lam = damp(eps_list[0])
        f_ba = sm.at_infinity(X_a, X_b, a, b, lam)
        g_ab = sm.at_infinity(X_b, X_a, b, a, lam)
        if debias:
            f_aa = sm.at_infinity(X_a, X_a, a, a, lam)
            g_bb = sm.at_infinity(X_b, X_b, b, b, lam)
        for eps in eps_list:
            lam = damp(eps)
            ft_ba = sm(eps, X_a, X_b, log_b, g_ab, damp=lam, old=f_ba)
            gt_ab = sm(eps, X_b, X_a, log_a, f_ba, damp=lam, old=g_ab)
            if debias:
                f_aa = sm(eps, X_a, X_a, log_a, f_aa, damp=lam, old=f_aa)
                g_bb = sm(eps, X_b, X_b, log_b, g_bb, damp=lam, old=g_bb)
            f_ba, g_ab = ft_ba, gt_ab
        if not want_grad:
            # last, non-averaged update (last_extrapolation=True, sinkhorn_ot.py:419-436)
            new_f = sm(eps, X_a, X_b, log_b, g_ab, damp=lam)
            new_g = sm(eps, X_b, X_a, log_a, f_ba, damp=lam)
            f_ba, g_ab = new_f, new_g
            if debias:
                f_aa = sm(eps, X_a, X_a, log_a, f_aa, damp=lam)
                g_bb = sm(eps, X_b, X_b, log_b, g_bb, damp=lam)
    if want_grad:
        # the same update, attached to the caller's clouds through the cost (both arguments of every softmin)
        xa, xb = live[0], live[1]
        last = lambda rows, cols, lw, pot: _LastUpdate.apply(rows, cols, lw, pot, eps, lam, sm.center)  # noqa: E731
        new_f, new_g = last(xa, xb, log_b, g_ab), last(xb, xa, log_a, f_ba)
        f_ba, g_ab = new_f, new_g
        if debias:
            f_aa, g_bb = last(xa, xa, log_a, f_aa), last(xb, xb, log_b, g_bb)
        a, b = live[2], live[3]
    if not debias:
        f_aa = g_bb = None
    return OTResultSample(X_a=X_a, X_b=X_b, a=a, b=b, reg=reg, unbalanced=rho, debias=bool(debias),
                          potentials=(f_aa, g_bb, g_ab, f_ba), softmin=sm)


def solve_sample_batch(*args, **kwargs):
    raise NotImplementedError("This function is not implemented yet.")  # as the reference (sample.py:405-431)


# ----------------------------------------------------------------------------------------------------
# results                                       ot/_ot_result.py:164-440, sample.py:444-641
# ----------------------------------------------------------------------------------------------------
class LinearOperator:
    """``x -> A @ x`` without the matrix (ot/_ot_result.py:9-160): ``op @ x``, ``op.T``, ``op.shape``."""

    def __init__(self, *, matmat, rmatmat, input_shape, output_shape):
        self._matmat, self._rmatmat = matmat, rmatmat
        self._input_shape, self._output_shape = tuple(input_shape), tuple(output_shape)

    def __matmul__(self, x):
        k = len(self._input_shape)
        if x.dim() < k or tuple(x.shape[:k]) != self._input_shape:
            raise ValueError(f"Expects an input of shape {self._input_shape} with, maybe, additional trailing "
                             f"dimensions, but found an array of shape {tuple(x.shape)}.")
        trailing = tuple(x.shape[k:])
        out = self._matmat(x.reshape(self._input_shape + (-1,)))
        return out.reshape(self._output_shape + trailing)

    @property
    def shape(self):
        return (math.prod(self._output_shape), math.prod(self._input_shape))

    def transpose(self):
        return LinearOperator(matmat=self._rmatmat, rmatmat=self._matmat, input_shape=self._output_shape,
                              output_shape=self._input_shape)

    @property
    def T(self):
        return self.transpose()

    def rescale(self, *, input_scaling, output_scaling):
        def matmat(s):
            return output_scaling[:, None] * (self._matmat(input_scaling[:, None] * s))

        def rmatmat(s):
            return input_scaling[:, None] * (self._rmatmat(output_scaling[:, None] * s))

        return LinearOperator(matmat=matmat, rmatmat=rmatmat, input_shape=self._input_shape,
                              output_shape=self._output_shape)


class OTResultSample:
    """Result of :func:`solve_sample`; lazily evaluated, cached attributes named like the reference's
    (``value``, ``potential_a/b/aa/bb``, ``marginal_a/b``, ``plan``, ``density``, ``plan_operator``,
    ``density_operator``; ``lazy_plan``, ``lazy_density``, ``a_to_b``, ``b_to_a`` are ``None`` as in the reference
    without KeOps)."""

    def __init__(self, *, X_a, X_b, a, b, reg, unbalanced, debias, potentials, softmin):
        self._X_a, self._X_b, self._a, self._b = X_a, X_b, a.detach(), b.detach()
        self._reg, self._unbalanced, self._debias = reg, unbalanced, debias
        self._live = (a, b) + tuple(potentials)  # possibly attached to the caller's graph: used by `value` only
        self._f_aa, self._g_bb, self._g_ab, self._f_ba = (None if t is None else t.detach() for t in potentials)
        self._sm = softmin

    # ---- dual potentials --------------------------------------------------------------------------
    @property
    def potential_a(self):
        return self._f_ba

    @property
    def potential_b(self):
        return self._g_ab

    @property
    def potential_aa(self):
        if self._f_aa is None:
            raise ValueError("The self-interaction potential `f_aa` is not defined. To fix this issue, run your OT "
                             "solver with `debias = True`.")
        return self._f_aa

    @property
    def potential_bb(self):
        if self._g_bb is None:
            raise ValueError("The self-interaction potential `g_bb` is not defined. To fix this issue, run your OT "
                             "solver with `debias = True`.")
        return self._g_bb

    # ---- value                                                   unbalanced_ot.py:25-185 ----------
    @cached_property
    def value(self):
        a, b, f_aa, g_bb, g_ab, f_ba = (None if t is None else t.double() for t in self._live)
        eps, rho = self._reg, self._unbalanced
        if rho is None:
            F, G = (f_ba - f_aa, g_ab - g_bb) if self._debias else (f_ba, g_ab)
        elif not self._debias:
            F = (rho + eps / 2 * b.sum()) - (rho + eps / 2) * torch.exp(-f_ba / rho)
            G = (rho + eps / 2 * a.sum()) - (rho + eps / 2) * torch.exp(-g_ab / rho)
        else:
            F = (rho + eps / 2) * (torch.exp(-f_aa / rho) - torch.exp(-f_ba / rho))
            G = (rho + eps / 2) * (torch.exp(-g_bb / rho) - torch.exp(-g_ab / rho))
        return ((a * F).sum() + (b * G).sum()).float()

    # ---- the plan as an operator: P s = exp((f - softmin(eps, log s, C, g)) / eps), s >= 0 -----------
    def _density_matmat(self, s, transpose):
        """(density @ s) for s of shape (M, V) [or density.T @ s, s of shape (N, V)], one pair of softmins per
        column: signed inputs are split into positive and negative parts, each a log-weight vector."""
        eps = self._reg
        rows, cols = (self._X_b, self._X_a) if transpose else (self._X_a, self._X_b)
        f, g = (self._g_ab, self._f_ba) if transpose else (self._f_ba, self._g_ab)
        out = []
        with torch.no_grad():
            for v in range(s.shape[1]):
                col = s[:, v].float().contiguous()
                acc = torch.zeros(rows.shape[0], dtype=torch.float32, device=rows.device)
                for sign in (1.0, -1.0):
                    part = (sign * col).clamp_min(0.0)
                    if bool((part > 0).any()):
                        smv = self._sm(eps, rows, cols, log_weights(part), g)
                        acc = acc + sign * torch.exp((f - smv) / eps)
                out.append(acc)
        return torch.stack(out, 1)

    @cached_property
    def density_operator(self):
        N, M = self._X_a.shape[0], self._X_b.shape[0]
        return LinearOperator(matmat=lambda s: self._density_matmat(s, False),
                              rmatmat=lambda s: self._density_matmat(s, True), input_shape=(M,), output_shape=(N,))

    @cached_property
    def plan_operator(self):
        return self.density_operator.rescale(input_scaling=self._b, output_scaling=self._a)

    @cached_property
    def marginal_a(self):
        return self._a * (self.density_operator @ self._b)

    @cached_property
    def marginal_b(self):
        return self._b * (self.density_operator.T @ self._a)

    # ---- dense accessors (small problems only) --------------------------------------------------------
    @cached_property
    def density(self):
        N, M = self._X_a.shape[0], self._X_b.shape[0]
        if N * M > _DENSE_LIMIT:
            raise MemoryError(f"a dense ({N}, {M}) density does not fit the budget of this accessor; use "
                              "density_operator / plan_operator, which never build the matrix")
        C = torch.cdist(self._X_a, self._X_b, compute_mode="donot_use_mm_for_euclid_dist") ** 2
        return torch.exp((self._f_ba[:, None] + self._g_ab[None, :] - C) / self._reg)

    @cached_property
    def plan(self):
        return self.density * self._a[:, None] * self._b[None, :]

    lazy_density = None
    lazy_plan = None
    a_to_b = None
    b_to_a = None

    def cache_clear(self):
        for k in ("value", "density_operator", "plan_operator", "marginal_a", "marginal_b", "density", "plan"):
            self.__dict__.pop(k, None)
