"""pykeops.torch — dense torch stand-ins for LazyTensor / generic_logsumexp (TEST INFRASTRUCTURE).

A ``LazyTensor`` here simply wraps the dense broadcast tensor ``(..., N|1, M|1, D)`` that the symbolic
pykeops object denotes; reductions materialise the full ``(..., N, M, D)`` array.  Conventions kept from
pykeops: the LAST axis is the vector dimension (``.sum(-1)`` keeps a trailing 1), the two axes before it are the
"i" and "j" axes, reductions over "j" (resp. "i") return ``(..., N, E)`` (resp. ``(..., M, E)``) torch tensors,
``K.ranges`` restricts the next reduction to the listed blocks, ``K @ v`` is the sum reduction over j.
"""
import re

import torch

from .cluster import ranges_to_mask, swap_axes


def _safe_sqrt(sq):
    """sqrt with KeOps' convention Rsqrt(0) = 0: value 0 and a ZERO (not infinite) gradient at 0."""
    pos = sq > 0
    return torch.where(pos, sq, torch.ones_like(sq)).sqrt() * pos.to(sq.dtype)


def _unwrap(other):
    return other.t_ if isinstance(other, LazyTensor) else other


class LazyTensor:
    def __init__(self, x, axis=None):
        if isinstance(x, (int, float)):
            x = torch.tensor(float(x))
        if axis is not None:  # Vi / Vj style construction from an (N, D) array
            x = x[:, None, :] if axis == 0 else x[None, :, :]
        self.t_ = x
        self.ranges = None

    # ---- shape helpers ----
    @property
    def shape(self):
        return tuple(self.t_.shape)

    @property
    def ndim(self):
        return self.t_.dim()

    def _wrap(self, t):
        out = LazyTensor(t)
        out.ranges = self.ranges
        return out

    def _lift(self, other):
        """Parameters given as plain tensors of shape (E,) broadcast against the vector axis."""
        return _unwrap(other)

    # ---- pointwise arithmetic ----
    def __add__(self, o):
        return self._wrap(self.t_ + self._lift(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self._wrap(self.t_ - self._lift(o))

    def __rsub__(self, o):
        return self._wrap(self._lift(o) - self.t_)

    def __mul__(self, o):
        return self._wrap(self.t_ * self._lift(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._wrap(self.t_ / self._lift(o))

    def __rtruediv__(self, o):
        return self._wrap(self._lift(o) / self.t_)

    def __neg__(self):
        return self._wrap(-self.t_)

    def __pow__(self, k):
        return self._wrap(self.t_**k)

    def abs(self):
        return self._wrap(self.t_.abs())

    def sqrt(self):
        return self._wrap(_safe_sqrt(self.t_))

    def exp(self):
        return self._wrap(self.t_.exp())

    def log(self):
        return self._wrap(self.t_.log())

    def square(self):
        return self._wrap(self.t_**2)

    # ---- axes ----
    def _axis(self, dim):
        """'v' for the vector axis, 'i' / 'j' for the reduction axes."""
        nd = self.t_.dim()
        if dim is None:
            raise ValueError("a reduction axis is required")
        if dim < 0:
            dim += nd
        if dim == nd - 1:
            return "v"
        if dim == nd - 2:
            return "j"
        if dim == nd - 3:
            return "i"
        raise ValueError(f"cannot reduce over batch axis {dim}")

    def _masked(self, fill):
        """Dense tensor with the pairs outside ``self.ranges`` replaced by ``fill`` (un-batched ranges only)."""
        t = self.t_
        if self.ranges is None:
            return t
        if t.dim() != 3:
            raise NotImplementedError("block-sparse ranges with batch dimensions")
        n_i, n_j = t.shape[0], t.shape[1]
        mask = ranges_to_mask(self.ranges, n_i, n_j, axis=1)
        return t.masked_fill(~mask[:, :, None], fill)

    def _expand(self, t):
        """Broadcast a (…,N|1,M|1,E) array to the full (…,N,M,E) problem (no-op when already full)."""
        return t

    def sum(self, dim=None, axis=None, **kw):
        dim = axis if dim is None else dim
        ax = self._axis(dim)
        if ax == "v":
            return self._wrap(self.t_.sum(-1, keepdim=True))
        t = self._masked(0.0)
        return t.sum(-2 if ax == "j" else -3)

    def logsumexp(self, dim=None, axis=None, weight=None, **kw):
        dim = axis if dim is None else dim
        ax = self._axis(dim)
        if ax == "v":
            raise NotImplementedError()
        t = self._masked(-float("inf"))
        return t.logsumexp(-2 if ax == "j" else -3)

    def max(self, dim=None, axis=None, **kw):
        dim = axis if dim is None else dim
        ax = self._axis(dim)
        t = self._masked(-float("inf"))
        return t.max(-2 if ax == "j" else -3).values

    def min(self, dim=None, axis=None, **kw):
        dim = axis if dim is None else dim
        ax = self._axis(dim)
        t = self._masked(float("inf"))
        return t.min(-2 if ax == "j" else -3).values

    def __matmul__(self, v):
        """Sum reduction over j of K_ij * v_j; v:(M,), (M,E) or batched (B,M,E)."""
        v = _unwrap(v)
        if v.dim() == 1:
            out = (self * LazyTensor(v[None, :, None])).sum(self.t_.dim() - 2)
            return out.squeeze(-1)
        vj = v.unsqueeze(-3)  # (…, 1, M, E)
        return (self * LazyTensor(vj)).sum(self.t_.dim() - 2)

    def t(self):
        out = LazyTensor(self.t_.transpose(-2, -3))
        if self.ranges is not None:
            out.ranges = swap_axes(self.ranges)
        return out

    @property
    def T(self):
        return self.t()


def Vi(x_or_ind, dim=None):
    if dim is None:
        return LazyTensor(x_or_ind, axis=0)
    raise NotImplementedError("symbolic Vi(ind, dim) variables are only used by disabled reference code")


def Vj(x_or_ind, dim=None):
    if dim is None:
        return LazyTensor(x_or_ind, axis=1)
    raise NotImplementedError("symbolic Vj(ind, dim) variables are only used by disabled reference code")


def Pm(x_or_ind, dim=None):
    if dim is None:
        return LazyTensor(x_or_ind)
    raise NotImplementedError("symbolic Pm(ind, dim) variables are only used by disabled reference code")


# ---- Genred-style string formulas -------------------------------------------------------------------------
_ALIAS = re.compile(r"\s*(\w+)\s*=\s*(Vi|Vj|Pm)\((?:(\d+)\s*,\s*)?(\d+)\)\s*")


def _formula_env():
    def SqDist(a, b):
        return ((a - b) ** 2).sum(-1, keepdim=True)

    def Norm2(a):
        return _safe_sqrt((a**2).sum(-1, keepdim=True))

    def SqNorm2(a):
        return (a**2).sum(-1, keepdim=True)

    def IntCst(k):
        return float(k)

    def Exp(a):
        return a.exp()

    def Sqrt(a):
        return _safe_sqrt(a)

    def Square(a):
        return a**2

    def Abs(a):
        return a.abs()

    return dict(SqDist=SqDist, Norm2=Norm2, SqNorm2=SqNorm2, IntCst=IntCst, Exp=Exp, Sqrt=Sqrt, Square=Square,
                Abs=Abs)


class _GenericReduction:
    """``generic_logsumexp(formula, out_alias, *arg_aliases)`` / ``generic_sum``: the output alias fixes the
    reduction axis (``Vi`` output: reduce over j), arguments are passed positionally in alias order."""

    def __init__(self, kind, formula, out_alias, *aliases, **kw):
        self.kind, self.formula = kind, formula
        m = _ALIAS.fullmatch(out_alias)
        self.out_cat = m.group(2)
        self.args = []
        for al in aliases:
            m = _ALIAS.fullmatch(al)
            if m is None:
                raise ValueError(f"cannot parse alias {al!r}")
            self.args.append((m.group(1), m.group(2), int(m.group(4))))

    def __call__(self, *tensors, ranges=None, **kw):
        env = _formula_env()
        n_i = n_j = None
        for (name, cat, dim), t in zip(self.args, tensors):
            if cat == "Vi":
                assert t.dim() == 2 and t.shape[1] == dim, (name, t.shape, dim)
                env[name] = t[:, None, :]
                n_i = t.shape[0]
            elif cat == "Vj":
                assert t.dim() == 2 and t.shape[1] == dim, (name, t.shape, dim)
                env[name] = t[None, :, :]
                n_j = t.shape[0]
            else:
                env[name] = t.view(1, 1, -1)
        val = eval(self.formula, {"__builtins__": {}}, env)  # (N, M, E) dense
        val = val.expand(n_i, n_j, val.shape[-1])
        axis = 1 if self.out_cat == "Vi" else 0
        if ranges is not None:
            mask = ranges_to_mask(ranges, n_i, n_j, axis=axis)
            val = val.masked_fill(~mask[:, :, None], -float("inf") if self.kind == "lse" else 0.0)
        return val.logsumexp(axis) if self.kind == "lse" else val.sum(axis)


def generic_logsumexp(formula, out_alias, *aliases, **kw):
    return _GenericReduction("lse", formula, out_alias, *aliases, **kw)


def generic_sum(formula, out_alias, *aliases, **kw):
    return _GenericReduction("sum", formula, out_alias, *aliases, **kw)
