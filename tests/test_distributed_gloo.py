"""World-size-2 ``gloo`` test (CPU) of the column-sharded engine's host logic: shard bounds, the
all_gather layout, the online log-sum-exp merge and the backward all_reduce.

The CUDA stage kernels cannot run here, so the test injects an oracle-backed stand-in for
``distributed.CudaStages`` (this is test infrastructure — the product default refuses CPU tensors).
The sharded SamplesLoss must reproduce the dense oracle: value, potentials and gradients.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


class OracleStages:
    """CPU restatement of the staged C-ABI calls, in the same (m, s) / partial-sum formats."""

    @staticmethod
    def _t(eps, x, y, h_a, h_b, h_scale_b, p):
        from oracle import geomloss_oracle as O

        h = h_a if h_b is None else h_a + h_scale_b * h_b
        # p carries B200OT_P_UNCLAMPED (0x100) for the pykeops cost convention
        C = O.keops_cost(x, y, p & 0xFF) if p & 0x100 else O.cost_matrix(x, y, p)
        return LOG2E * (h[None, :] - C / eps)

    def softmin_shard(self, eps, x, y, h_a, h_b, h_scale_b, p, center):
        t = self._t(eps, x.double(), y.double(), h_a.double(), None if h_b is None else h_b.double(), h_scale_b, p)
        m = t.max(1).values
        s = torch.exp2(t - m[:, None]).sum(1)
        return torch.stack([m, s], 1).float()

    def empty_shard(self, N, dev):
        out = torch.zeros(N, 2)
        out[:, 0] = -1.0e30
        return out

    def softmin_finalize(self, parts, eps, out_old, alpha_old, beta, want_lse2):
        parts = parts.double()
        mm = parts[..., 0].max(0).values
        ss = (parts[..., 1] * torch.exp2(parts[..., 0] - mm)).sum(0)
        lse2 = mm + torch.log2(ss)
        out = -eps * LN2 * lse2 * beta
        if out_old is not None:
            out = out + alpha_old * out_old.double()
        return out.float(), (lse2.float() if want_lse2 else None)

    def softmin_bwd_shard(self, eps, x, y, h_a, h_b, h_scale_b, p, center, lse2):
        xd, yd = x.double(), y.double()
        t = self._t(eps, xd, yd, h_a.double(), None if h_b is None else h_b.double(), h_scale_b, p)
        w = torch.exp2(t - lse2.double()[:, None])
        return self._bwd_sums(w, xd, yd, p)

    @staticmethod
    def _bwd_sums(w, xd, yd, p):
        if p & 0xFF == 2:
            vec = w @ yd  # un-scaled, un-centred coordinates: finalize below matches
        else:
            clamp = 1e-30 if p & 0x100 else 1e-8
            diff = xd[:, None, :] - yd[None, :, :]
            q = (diff**2).sum(-1)
            unit = torch.where(q[..., None] < clamp, torch.zeros_like(diff), diff / q.clamp_min(clamp).sqrt()[..., None])
            vec = (w[..., None] * unit).sum(1)
        return torch.cat([w.sum(1, keepdim=True), vec], 1).float()

    def softmin_bwd_finalize(self, sums, eps, x, center, grad_out, p):
        sums = sums.double()
        sw = sums[:, :1]
        g = (x.double() - sums[:, 1:] / sw) if p & 0xFF == 2 else sums[:, 1:] / sw
        return (grad_out.double()[:, None] * g).float()

    # -- ranges-mode stand-ins: dense evaluation, masked with the (row, column) pairs the descriptors list --
    ranges_variant = 1  # small tiles (128-row segments, 256-column pieces): several segments per cluster here

    @staticmethod
    def _pair_mask(N, M, prob):
        """Point-level mask of a RangesProblem: what b200ot_softmin_partial_ranges would visit."""
        src = prob.layout.src.tolist()
        mask = torch.zeros(N, M, dtype=torch.bool)
        pieces = prob.pieces.tolist()
        for r0, nr, p0, p1 in prob.seg.tolist():
            for c0, nc in pieces[p0:p1]:
                cols = [j for j in src[c0:c0 + nc] if j >= 0]
                mask[r0:r0 + nr, cols] = True
        return mask

    def softmin_sparse_shard(self, eps, x, y, h_a, h_b, h_scale_b, p, center, prob):
        t = self._t(eps, x.double(), y.double(), h_a.double(), None if h_b is None else h_b.double(), h_scale_b, p)
        t = t.masked_fill(~self._pair_mask(x.shape[0], y.shape[0], prob), -1.0e30)
        m = t.max(1).values
        s = torch.exp2(t - m[:, None]).sum(1)
        s = torch.where(m <= -1.0e29, torch.zeros_like(s), s)  # a row with no listed piece: neutral partial
        return torch.stack([m, s], 1).float()

    def softmin_bwd_sparse_shard(self, eps, x, y, h_a, h_b, h_scale_b, p, center, lse2, prob):
        xd, yd = x.double(), y.double()
        t = self._t(eps, xd, yd, h_a.double(), None if h_b is None else h_b.double(), h_scale_b, p)
        w = torch.exp2(t - lse2.double()[:, None]) * self._pair_mask(x.shape[0], y.shape[0], prob)
        return self._bwd_sums(w, xd, yd, p)

    @staticmethod
    def _kmat(kind, x, y, blur):
        from oracle import geomloss_oracle as O

        if isinstance(kind, str):
            return O.kernel_matrix(kind, x, y, blur)
        name = ("gaussian", "laplacian", "energy")[kind & 0xFF]
        return O.keops_kernel_matrix(name, x, y, blur) if kind & 0x100 else O.kernel_matrix(name, x, y, blur)

    def conv_shard(self, kind, x, y, w, blur, center):
        return (self._kmat(kind, x.double(), y.double(), blur) @ w.double()).float()

    def conv_grad_shard(self, kind, x, y, w, blur, grad_out, center):
        from oracle import geomloss_oracle as O

        with torch.enable_grad():  # called from inside an autograd backward (grad mode is off there)
            xr = x.detach().double().requires_grad_(True)
            out = self._kmat(kind, xr, y.detach().double(), blur) @ w.detach().double()
            (g,) = torch.autograd.grad(out, xr, grad_out.detach().double())
        return g.float()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geomloss_b200 import SamplesLoss
        from geomloss_b200.distributed import ColumnShardedEngine, shard_bounds
        from oracle import geomloss_oracle as O

        torch.set_num_threads(2)
        g = torch.Generator().manual_seed(0)
        x, y = torch.rand(151, 3, generator=g), torch.rand(97, 3, generator=g)  # odd sizes: uneven shards
        a = torch.rand(151, generator=g)
        b = torch.rand(97, generator=g)
        a, b = a / a.sum(), b / b.sum()
        out = {}
        for tag, kw in (("bal", dict(p=2, blur=0.1)), ("unb_p1", dict(p=1, blur=0.1, reach=0.4))):
            eng = ColumnShardedEngine(stages=OracleStages())
            assert shard_bounds(97, 0, 2) == (0, 48) and shard_bounds(97, 1, 2) == (48, 97)
            xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            L = eng.attach(SamplesLoss("sinkhorn", **kw))
            val = L(a, xg, b, yg)
            gx, gy = torch.autograd.grad(val, [xg, yg])
            # fp64 dense oracle (the stand-in stages above also work in fp64)
            xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
            ref = O.samples_loss(a.double(), xr, b.double(), yr, loss="sinkhorn", **kw)
            rx, ry = torch.autograd.grad(ref, [xr, yr])
            F, G = eng.attach(SamplesLoss("sinkhorn", potentials=True, **kw))(a, x, b, y)
            Fr, Gr = O.samples_loss(a.double(), x.double(), b.double(), y.double(), loss="sinkhorn",
                                    potentials=True, **kw)
            out[tag] = dict(val=val.item(), ref=ref.item(), gx=(gx - rx).abs().max().item(),
                            gy=(gy - ry).abs().max().item(), gscale=rx.abs().max().item(),
                            F=(F - Fr).abs().max().item(), G=(G - Gr).abs().max().item(),
                            collectives=eng.collectives)
        # kernel MMDs: sharded matvecs (all_reduce) and their three gradients (all_reduce / all_gather)
        for kind in ("gaussian", "energy"):
            eng = ColumnShardedEngine(stages=OracleStages())
            ag, xg = a.clone().requires_grad_(True), x.clone().requires_grad_(True)
            bg, yg = b.clone().requires_grad_(True), y.clone().requires_grad_(True)
            val = eng.attach(SamplesLoss(kind, blur=0.3))(ag, xg, bg, yg)
            grads = torch.autograd.grad(val, [ag, xg, bg, yg])
            ar, xr = a.double().requires_grad_(True), x.double().requires_grad_(True)
            br, yr = b.double().requires_grad_(True), y.double().requires_grad_(True)
            ref = O.samples_loss(ar, xr, br, yr, loss=kind, blur=0.3)
            rgrads = torch.autograd.grad(ref, [ar, xr, br, yr])
            out["mmd_" + kind] = dict(val=val.item(), ref=ref.item(),
                                      gerr=max((g.double() - r).abs().max().item() / max(r.abs().max().item(), 1e-12)
                                               for g, r in zip(grads, rgrads)), collectives=eng.collectives)
        # two-scale (multiscale) Sinkhorn: coarse phase replicated, block-sparse fine phase column-sharded by tiles
        gm = torch.Generator().manual_seed(5)
        xm, ym = torch.rand(700, 3, generator=gm), torch.rand(610, 3, generator=gm) * 0.8 + 0.1
        am, bm = torch.rand(700, generator=gm) + 0.5, torch.rand(610, generator=gm) + 0.5
        am, bm = am / am.sum(), bm / bm.sum()
        solo_groups = [dist.new_group([r]) for r in range(world)]  # every rank alone: the un-sharded run
        for tag, kw, truncate in (("ms_exact", dict(p=2, blur=0.05), None), ("ms_trunc", dict(p=2, blur=0.05), 5),
                                  ("ms_lastjump", dict(p=2, blur=0.4, reach=0.5), 5)):
            eng = ColumnShardedEngine(stages=OracleStages())
            L = eng.attach(SamplesLoss("sinkhorn", backend="multiscale", cluster_scale=0.2, truncate=truncate, **kw))
            xg = xm.clone().requires_grad_(True)
            val = L(am, xg, bm, ym)
            (gx,) = torch.autograd.grad(val, xg)
            ref = O.sinkhorn_multiscale_dense(am.double(), xm.double(), bm.double(), ym.double(), cluster_scale=0.2,
                                              truncate=truncate, **kw)
            # gradient contract (last step only, detached columns) is pinned on the GPU; here: sharded == un-sharded
            solo = ColumnShardedEngine(group=solo_groups[rank], stages=OracleStages())
            xs = xm.clone().requires_grad_(True)
            (rx,) = torch.autograd.grad(solo.attach(SamplesLoss("sinkhorn", backend="multiscale", cluster_scale=0.2,
                                                                truncate=truncate, **kw))(am, xs, bm, ym), xs)
            F, G = eng.attach(SamplesLoss("sinkhorn", backend="multiscale", cluster_scale=0.2, truncate=truncate,
                                          potentials=True, **kw))(am, xm, bm, ym)
            Fr, Gr = O.sinkhorn_multiscale_dense(am.double(), xm.double(), bm.double(), ym.double(),
                                                 cluster_scale=0.2, truncate=truncate, potentials=True, **kw)
            out[tag] = dict(val=val.item(), ref=ref.item(), gx=(gx - rx).abs().max().item(),
                            gscale=rx.abs().max().item(), F=(F - Fr).abs().max().item(),
                            G=(G - Gr).abs().max().item(), collectives=eng.collectives)
        # every rank must end with the same numbers (replicated state)
        vals = [None] * world
        dist.all_gather_object(vals, out["bal"]["val"])
        out["replicated"] = all(v == vals[0] for v in vals)
        if rank == 0:
            results.put(out)
    except Exception as exc:  # surface the failure instead of letting the parent time out
        import traceback

        results.put({"error": f"rank {rank}: {exc!r}\n{traceback.format_exc()}"})
        raise
    finally:
        dist.destroy_process_group()


def test_column_sharded_engine_world2():
    world = 2
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    out = results.get(timeout=240)
    assert "error" not in out, out.get("error")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["replicated"]
    for kind in ("gaussian", "energy"):
        r = out["mmd_" + kind]
        assert abs(r["val"] - r["ref"]) <= 1e-5 * abs(r["ref"]) + 1e-9, r
        assert r["gerr"] < 1e-4 and r["collectives"] > 0, r
    for tag in ("ms_exact", "ms_trunc", "ms_lastjump"):
        r = out[tag]
        assert abs(r["val"] - r["ref"]) <= 5e-5 * abs(r["ref"]), (tag, r)
        assert r["gx"] <= 1e-5 * r["gscale"] and r["gscale"] > 0, (tag, r)
        assert r["F"] < 2e-5 and r["G"] < 2e-5, (tag, r)
        assert r["collectives"] > 0
    for tag in ("bal", "unb_p1"):
        r = out[tag]
        assert abs(r["val"] - r["ref"]) <= 2e-5 * abs(r["ref"]), r
        assert r["gx"] <= 2e-4 * r["gscale"] and r["gy"] <= 2e-4 * r["gscale"], r
        assert r["F"] < 5e-6 and r["G"] < 5e-6, r
        # 4 softmins x (init + n_eps + final) forward collectives + 4 backward all_reduces
        assert r["collectives"] > 0 and r["collectives"] % 4 == 0


def test_shard_bounds_cover_and_balance():
    from geomloss_b200.distributed import shard_bounds

    for M in (1, 7, 8, 97, 10**6, 10**7 + 3):
        for W in (1, 2, 3, 8):
            edges = [shard_bounds(M, r, W) for r in range(W)]
            assert edges[0][0] == 0 and edges[-1][1] == M
            assert all(edges[i][1] == edges[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_engine_requires_initialised_process_group():
    from geomloss_b200.distributed import ColumnShardedEngine

    if dist.is_initialized():
        pytest.skip("a process group is active in this interpreter")
    with pytest.raises(RuntimeError):
        ColumnShardedEngine()
