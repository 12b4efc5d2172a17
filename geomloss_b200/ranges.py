"""Ranges mode of the reductions: block-sparse and batched problems in ONE launch group.

Host side of the reference's ``ranges=`` machinery — pykeops' ``from_matrix`` / ``cluster_ranges`` tuples consumed by
``softmin_multiscale`` (src/geomloss/_legacy/sinkhorn_samples.py:445-450, built by ``kernel_truncation`` :493-530) and
``kernel_multiscale`` (src/geomloss/_legacy/kernel_samples.py:246-271) — and of the batched LazyTensor reduction
(``softmin_online_lazytensor``, sinkhorn_samples.py:229-290), which is the block-diagonal special case.

A problem is described to libb200ot.so (include/b200ot.h, "ranges mode") by
  * a ``ColumnLayout``: the column cloud is grouped (clusters of a label-sorted cloud, or batch elements); every
    group is padded to a multiple of 16 column SLOTS so that it starts on a chunk boundary of the kernels; the pack
    kernel gathers ``slot -> column`` through ``src`` (padding slots are neutral columns);
  * segments ``(row_start, row_count, piece_begin, piece_end)``: runs of rows of ONE row group, at most one CTA tall;
  * pieces ``(col_start, col_count)``: runs of consecutive kept column groups, cut to at most one tile.
The kernels then visit exactly the (row group, column group) pairs of the boolean ``keep`` matrix — the same blocks
as the reference, not a tile-level superset.  Everything below is vectorised torch on the device (no Python loop
over clusters); the only host syncs are the sizes of the descriptor arrays.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops

BIG, SMALL = 0, 1


def shape(variant):
    r, c, al = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
    _lib.lib().b200ot_ranges_shape(int(variant), ctypes.byref(r), ctypes.byref(c), ctypes.byref(al))
    return r.value, c.value, al.value


def _excl_cumsum(t):
    out = torch.zeros(t.numel() + 1, dtype=torch.int64, device=t.device)
    out[1:] = t.cumsum(0)
    return out


class ColumnLayout:
    """Slot layout of a grouped column cloud: ``counts[g]`` consecutive columns belong to group g."""

    def __init__(self, counts, align=16):
        counts = counts.to(torch.int64)
        dev = counts.device
        self.counts = counts
        self.padded = (counts + align - 1) // align * align
        self.col_start = _excl_cumsum(counts)
        self.slot_start = _excl_cumsum(self.padded)
        self.n_cols = int(self.col_start[-1].item())
        self.n_slots = int(self.slot_start[-1].item())
        grp = torch.repeat_interleave(torch.arange(counts.numel(), device=dev), counts)
        cols = torch.arange(self.n_cols, device=dev)
        slot_of_col = self.slot_start[grp] + (cols - self.col_start[grp])
        src = torch.full((max(self.n_slots, 1),), -1, dtype=torch.int32, device=dev)
        src[slot_of_col] = cols.to(torch.int32)
        self.src = src.contiguous()


class RangesProblem:
    """Descriptor arrays of one reduction problem (rows x grouped columns restricted to ``keep``)."""

    def __init__(self, seg, pieces, layout, variant, n_rows, density):
        self.seg, self.pieces, self.layout, self.variant = seg, pieces, layout, variant
        self.n_seg = int(seg.shape[0])
        self.n_rows = n_rows
        self.density = density  # kept (row, column) pairs / all pairs


def pick_variant(row_counts, col_counts):
    """Big tiles (512-row segments, 1024-column pieces) pay off once groups hold a few hundred points."""
    rows = float(row_counts.double().mean().item())
    cols = float(col_counts.double().mean().item())
    return BIG if (rows >= 192 and cols >= 192) else SMALL


def build_problem(keep, row_counts, layout, variant=None, rank=0, world=1):
    """``keep``: (R, C) bool over (row groups, column groups), or None for all pairs.  ``row_counts``: (R,) rows per
    row group (consecutive).  ``world > 1``: column-sharded run — this rank keeps a contiguous slab of column groups,
    cut so that every rank reduces about the same number of pairs."""
    dev = row_counts.device
    row_counts = row_counts.to(torch.int64)
    R, C = row_counts.numel(), layout.counts.numel()
    if variant is None:
        variant = pick_variant(row_counts, layout.counts)
    max_rows, max_cols, _ = shape(variant)
    if keep is None:
        keep = torch.ones(R, C, dtype=torch.bool, device=dev)
    keep = keep.to(torch.bool)
    pairs_all = float(row_counts.sum().item()) * float(layout.counts.sum().item())
    w_pairs = (keep.double() * row_counts.double()[:, None] * layout.counts.double()[None, :])
    density = float(w_pairs.sum().item()) / max(pairs_all, 1.0)
    if world > 1:
        load = w_pairs.sum(0).cumsum(0)
        total = load[-1]
        cuts = torch.searchsorted(load, total * torch.arange(1, world, device=dev, dtype=torch.float64) / world)
        cuts = [0] + [min(int(c) + 1, C) for c in cuts.tolist()] + [C]
        lo, hi = cuts[rank], max(cuts[rank], cuts[rank + 1])
        keep = keep.clone()
        keep[:, :lo] = False
        keep[:, hi:] = False
    # maximal runs of consecutive kept column groups, row group by row group
    false_col = torch.zeros(R, 1, dtype=torch.bool, device=dev)
    start = keep & ~torch.cat([false_col, keep[:, :-1]], dim=1)
    end = keep & ~torch.cat([keep[:, 1:], false_col], dim=1)
    s_idx, e_idx = start.nonzero(), end.nonzero()  # row-major order: the k-th start pairs with the k-th end
    run_row = s_idx[:, 0]
    run_lo = layout.slot_start[s_idx[:, 1]]
    run_hi = layout.slot_start[e_idx[:, 1] + 1]
    n_piece = (run_hi - run_lo + max_cols - 1) // max_cols
    n_runs = run_row.numel()
    piece_run = torch.repeat_interleave(torch.arange(n_runs, device=dev), n_piece)
    first = n_piece.cumsum(0) - n_piece
    k_in = torch.arange(piece_run.numel(), device=dev) - first[piece_run]
    col_start = run_lo[piece_run] + k_in * max_cols
    col_count = torch.minimum(torch.full_like(col_start, max_cols), run_hi[piece_run] - col_start)
    pieces = torch.stack([col_start, col_count], dim=1).to(torch.int32).contiguous()
    per_row = torch.zeros(R, dtype=torch.int64, device=dev).index_add_(0, run_row, n_piece)
    piece_ptr = _excl_cumsum(per_row)
    # segments: every row group is cut into CTA-sized runs of rows sharing its piece list
    row_off = _excl_cumsum(row_counts)
    n_seg_r = (row_counts + max_rows - 1) // max_rows
    seg_grp = torch.repeat_interleave(torch.arange(R, device=dev), n_seg_r)
    k_seg = torch.arange(seg_grp.numel(), device=dev) - (n_seg_r.cumsum(0) - n_seg_r)[seg_grp]
    row_start = row_off[seg_grp] + k_seg * max_rows
    row_count = torch.minimum(torch.full_like(row_start, max_rows), row_off[seg_grp + 1] - row_start)
    seg = torch.stack([row_start, row_count, piece_ptr[seg_grp], piece_ptr[seg_grp + 1]], dim=1)
    # heaviest segments first: shortens the tail of the last wave
    work = (seg[:, 3] - seg[:, 2]) * max_cols
    seg = seg[torch.argsort(work, descending=True, stable=True)].to(torch.int32).contiguous()
    if pieces.shape[0] == 0:
        pieces = torch.zeros(1, 2, dtype=torch.int32, device=dev)
    return RangesProblem(seg, pieces, layout, variant, int(row_off[-1].item()), density)


def batch_problem(B, N, M, device, variant=None):
    """B independent (N rows) x (M columns) problems stacked along the row / column axes (block-diagonal keep)."""
    layout = ColumnLayout(torch.full((B,), M, dtype=torch.int64, device=device))
    rows = torch.full((B,), N, dtype=torch.int64, device=device)
    if variant is None:
        variant = BIG if (N >= 2048 and M >= 2048) else SMALL
    return build_problem(torch.eye(B, dtype=torch.bool, device=device), rows, layout, variant=variant)


# ------------------------------------------------------------------------------------------------------
# softmin
# ------------------------------------------------------------------------------------------------------
def softmin_shard_ranges(eps, x, y, h_a, h_b, h_scale_b, prob, *, p=2, center=None, part=None):
    """pack (gather) + partial reduction: the (N, 2) (m, s) partials of this problem's pieces."""
    x, y, h_a, h_b, center = (ops._f32c(t, n) for t, n in ((x, "x"), (y, "y"), (h_a, "h_a"), (h_b, "h_b"),
                                                           (center, "center")))
    N, D = x.shape
    if N != prob.n_rows or y.shape[0] != prob.layout.n_cols:
        raise ValueError("point clouds do not match the ranges problem they are reduced with")
    dev = x.device
    L = _lib.lib()
    lay = prob.layout
    with torch.cuda.device(dev):
        if part is None:
            part = torch.empty(N, 2, dtype=torch.float32, device=dev)
        cols = ops._scratch(L.b200ot_packed_cols_floats(lay.n_slots, D, 1) * 4, dev, "ranges_cols")
        st = ops._stream(dev)
        _lib.check(L.b200ot_softmin_pack_gather(ops._ptr(y), ops._ptr(h_a), ops._ptr(h_b), float(h_scale_b),
                                                ops._ptr(center), ops._ptr(lay.src), lay.n_slots, D, int(p),
                                                float(eps), ops._ptr(cols), st), "b200ot_softmin_pack_gather")
        _lib.check(L.b200ot_softmin_partial_ranges(ops._ptr(x), ops._ptr(center), ops._ptr(cols), ops._ptr(prob.seg),
                                                   prob.n_seg, ops._ptr(prob.pieces), ops._ptr(part), N, D, int(p),
                                                   float(eps), prob.variant, st), "b200ot_softmin_partial_ranges")
    ops.count_launches(2)
    return part


def softmin_finalize(parts, n_part, eps, out_old, alpha_old, beta, want_lse2, out=None):
    N = parts.shape[-2]
    dev = parts.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(N, dtype=torch.float32, device=dev)
        lse2 = torch.empty(N, dtype=torch.float32, device=dev) if want_lse2 else None
        _lib.check(L.b200ot_softmin_finalize(ops._ptr(parts), int(n_part), ops._ptr(ops._f32c(out_old, "out_old")),
                                             float(alpha_old), float(beta), ops._ptr(out), ops._ptr(lse2), N,
                                             float(eps), ops._stream(dev)), "b200ot_softmin_finalize")
    ops.count_launches(1)
    return out, lse2


def softmin_ranges_raw(eps, x, y, h_a, h_b, h_scale_b, prob, *, p=2, center=None, out_old=None, alpha_old=0.0,
                       beta=1.0, want_lse2=False, out=None):
    part = softmin_shard_ranges(eps, x, y, h_a, h_b, h_scale_b, prob, p=p, center=center)
    return softmin_finalize(part, 1, eps, out_old, alpha_old, beta, want_lse2, out=out)


def softmin_bwd_shard_ranges(eps, x, y, h_a, h_b, h_scale_b, prob, lse2, *, p=2, center=None):
    """(N, D+1) row-gradient sums over this problem's pieces."""
    x, y, h_a, h_b, center, lse2 = (ops._f32c(t, n) for t, n in ((x, "x"), (y, "y"), (h_a, "h_a"), (h_b, "h_b"),
                                                                 (center, "center"), (lse2, "lse2")))
    N, D = x.shape
    dev = x.device
    L = _lib.lib()
    lay = prob.layout
    with torch.cuda.device(dev):
        sums = torch.zeros(N, D + 1, dtype=torch.float32, device=dev)
        cols = ops._scratch(L.b200ot_packed_cols_floats(lay.n_slots, D, 1) * 4, dev, "ranges_cols")
        st = ops._stream(dev)
        _lib.check(L.b200ot_softmin_pack_gather(ops._ptr(y), ops._ptr(h_a), ops._ptr(h_b), float(h_scale_b),
                                                ops._ptr(center), ops._ptr(lay.src), lay.n_slots, D, int(p),
                                                float(eps), ops._ptr(cols), st), "b200ot_softmin_pack_gather")
        _lib.check(L.b200ot_softmin_bwd_partial_ranges(ops._ptr(x), ops._ptr(center), ops._ptr(cols), ops._ptr(lse2),
                                                       ops._ptr(prob.seg), prob.n_seg, ops._ptr(prob.pieces),
                                                       ops._ptr(sums), N, D, int(p), float(eps), prob.variant, st),
                   "b200ot_softmin_bwd_partial_ranges")
    ops.count_launches(2)
    return sums


def softmin_bwd_finalize(sums, eps, x, center, grad_out, p):
    x, center, grad_out = ops._f32c(x, "x"), ops._f32c(center, "center"), ops._f32c(grad_out, "grad_out")
    N, D = x.shape
    dev = x.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        gx = torch.empty_like(x)
        _lib.check(L.b200ot_softmin_bwd_finalize(ops._ptr(sums), 1, ops._ptr(x), ops._ptr(center),
                                                 ops._ptr(grad_out), ops._ptr(gx), N, D, int(p), float(eps),
                                                 ops._stream(dev)), "b200ot_softmin_bwd_finalize")
    ops.count_launches(1)
    return gx


class _RangesSoftmin(torch.autograd.Function):
    """Ranges-mode softmin with the reference's autograd contract (gradient to the row cloud only)."""

    @staticmethod
    def forward(ctx, x, y, h_a, h_b, h_scale_b, eps, p, center, scale_out, prob):
        need = ctx.needs_input_grad[0]
        out, lse2 = softmin_ranges_raw(eps, x, y, h_a, h_b, h_scale_b, prob, p=p, center=center, beta=scale_out,
                                       want_lse2=need)
        if need:
            ctx.save_for_backward(x, y, h_a, h_b if h_b is not None else h_a, center if center is not None else h_a,
                                  lse2)
            ctx.meta = (float(h_scale_b), float(eps), int(p), float(scale_out), prob, h_b is not None,
                        center is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, y, h_a, h_b, center, lse2 = ctx.saved_tensors
        h_scale_b, eps, p, scale_out, prob, has_hb, has_center = ctx.meta
        center = center if has_center else None
        sums = softmin_bwd_shard_ranges(eps, x, y, h_a, h_b if has_hb else None, h_scale_b, prob, lse2, p=p,
                                        center=center)
        gx = softmin_bwd_finalize(sums, eps, x, center, (grad_out * scale_out).contiguous(), p)
        return gx, None, None, None, None, None, None, None, None, None


def softmin_ranges(eps, x, y, h_a, h_b, h_scale_b, prob, *, p=2, center=None, scale_out=1.0):
    return _RangesSoftmin.apply(x, y.detach(), h_a.detach(), None if h_b is None else h_b.detach(), h_scale_b, eps, p,
                                center, scale_out, prob)


# ------------------------------------------------------------------------------------------------------
# kernel convolutions
# ------------------------------------------------------------------------------------------------------
def _conv_width(kind_id, D, backward):
    if not backward:
        return 1
    return D + 1 if (kind_id & 0xFF) == 0 else D


def conv_shard_ranges(kind_id, x, y, w, blur, prob, *, center=None, backward=False):
    """(N,) sums (forward) or (N, width) row-gradient sums over this problem's pieces."""
    x, y, w, center = ops._f32c(x, "x"), ops._f32c(y, "y"), ops._f32c(w, "w"), ops._f32c(center, "center")
    N, D = x.shape
    if N != prob.n_rows or y.shape[0] != prob.layout.n_cols:
        raise ValueError("point clouds do not match the ranges problem they are reduced with")
    dev = x.device
    L = _lib.lib()
    lay = prob.layout
    width = _conv_width(kind_id, D, backward)
    with torch.cuda.device(dev):
        part = torch.zeros(N, width, dtype=torch.float32, device=dev)
        cols = ops._scratch(L.b200ot_packed_cols_floats(lay.n_slots, D, 2) * 4, dev, "ranges_cols")
        st = ops._stream(dev)
        _lib.check(L.b200ot_kernel_conv_pack_gather(ops._ptr(y), ops._ptr(w), ops._ptr(center), ops._ptr(lay.src),
                                                    lay.n_slots, D, int(kind_id), float(blur), ops._ptr(cols), st),
                   "b200ot_kernel_conv_pack_gather")
        _lib.check(L.b200ot_kernel_conv_partial_ranges(ops._ptr(x), ops._ptr(center), ops._ptr(cols),
                                                       ops._ptr(prob.seg), prob.n_seg, ops._ptr(prob.pieces),
                                                       ops._ptr(part), N, D, int(kind_id), float(blur),
                                                       1 if backward else 0, prob.variant, st),
                   "b200ot_kernel_conv_partial_ranges")
    ops.count_launches(2)
    return part


def conv_finalize(kind_id, part):
    N = part.shape[0]
    dev = part.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        out = torch.empty(N, dtype=torch.float32, device=dev)
        _lib.check(L.b200ot_kernel_conv_finalize(ops._ptr(part), 1, ops._ptr(out), N, int(kind_id), ops._stream(dev)),
                   "b200ot_kernel_conv_finalize")
    ops.count_launches(1)
    return out


def conv_bwd_finalize(kind_id, part, x, center, grad_out, blur):
    x, center, grad_out = ops._f32c(x, "x"), ops._f32c(center, "center"), ops._f32c(grad_out, "grad_out")
    N, D = x.shape
    dev = x.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        gx = torch.empty_like(x)
        _lib.check(L.b200ot_kernel_conv_bwd_finalize(ops._ptr(part), 1, ops._ptr(x), ops._ptr(center),
                                                     ops._ptr(grad_out), ops._ptr(gx), N, D, int(kind_id),
                                                     float(blur), ops._stream(dev)), "b200ot_kernel_conv_bwd_finalize")
    ops.count_launches(1)
    return gx


class _RangesConv(torch.autograd.Function):
    """out = (K(x, y) restricted to prob) @ w, differentiable w.r.t. x, y and w; ``prob_t`` is the transposed problem
    (rows y, columns x) that serves the y- and w-gradients (kernels are symmetric)."""

    @staticmethod
    def forward(ctx, x, y, w, kind_id, blur, center, prob, prob_t):
        out = conv_finalize(kind_id, conv_shard_ranges(kind_id, x, y, w, blur, prob, center=center))
        ctx.save_for_backward(x, y, w, center if center is not None else w)
        ctx.meta = (kind_id, float(blur), center is not None, prob, prob_t)
        return out

    @staticmethod
    def backward(ctx, go):
        x, y, w, center = ctx.saved_tensors
        kind_id, blur, has_center, prob, prob_t = ctx.meta
        center = center if has_center else None
        go = go.contiguous()
        gx = gy = gw = None
        if ctx.needs_input_grad[0]:
            part = conv_shard_ranges(kind_id, x, y, w, blur, prob, center=center, backward=True)
            gx = conv_bwd_finalize(kind_id, part, x, center, go, blur)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            if prob_t is None:
                raise RuntimeError("the transposed ranges problem is needed for gradients w.r.t. y / w")
        if ctx.needs_input_grad[1]:
            part = conv_shard_ranges(kind_id, y, x, go, blur, prob_t, center=center, backward=True)
            gy = conv_bwd_finalize(kind_id, part, y, center, w, blur)
        if ctx.needs_input_grad[2]:
            gw = conv_finalize(kind_id, conv_shard_ranges(kind_id, y, x, go, blur, prob_t, center=center))
        return gx, gy, gw, None, None, None, None, None


def kernel_conv_ranges(kind_id, x, y, w, blur, prob, prob_t=None, *, center=None):
    return _RangesConv.apply(x, y, w, kind_id, blur, center, prob, prob_t)
