"""Multi-GPU (NCCL) test of the column-sharded engine: needs >= 2 CUDA devices, otherwise skipped.

Spawns one process per GPU, runs the Sinkhorn loss with the columns of every softmin sharded across the
ranks and checks (a) every rank ends with the same value / gradients, (b) they agree with the
single-GPU engine and the CPU oracle.
"""
import os
import socket
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from geomloss_b200 import SamplesLoss
        from geomloss_b200.distributed import ColumnShardedEngine

        g = torch.Generator().manual_seed(0)
        x = torch.rand(7001, 3, generator=g).to(dev)
        y = torch.rand(5003, 3, generator=g).to(dev)
        out = {}
        for tag, kw in (("p2", dict(p=2, blur=0.05)), ("p1_unb", dict(p=1, blur=0.05, reach=0.5))):
            xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            single = SamplesLoss("sinkhorn", **kw)(xs, ys)
            gxs, gys = torch.autograd.grad(single, [xs, ys])
            eng = ColumnShardedEngine()
            xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            val = eng.attach(SamplesLoss("sinkhorn", **kw))(xg, yg)
            gx, gy = torch.autograd.grad(val, [xg, yg])
            vals = [None] * world
            dist.all_gather_object(vals, (val.item(), gx.abs().sum().item()))
            out[tag] = dict(val=val.item(), single=single.item(),
                            gx=(gx - gxs).abs().max().item() / gxs.abs().max().item(),
                            gy=(gy - gys).abs().max().item() / gys.abs().max().item(),
                            replicated=all(v == vals[0] for v in vals), collectives=eng.collectives)
        # D = 16: shards go through the tensor-core entries (b200ot_softmin_fwd -> (lse2, 1) partials,
        # b200ot_softmin_bwd_sums) — ADVICE r01: the staged CUDA-core calls reject D > 8
        gh = torch.Generator().manual_seed(1)
        xh, yh = torch.rand(3001, 16, generator=gh).to(dev), torch.rand(2503, 16, generator=gh).to(dev)
        for tag, loss, kw in (("d16_sinkhorn", "sinkhorn", dict(p=2, blur=0.5, scaling=0.6)),
                              ("d16_gaussian", "gaussian", dict(blur=1.5))):
            xs, ys = xh.clone().requires_grad_(True), yh.clone().requires_grad_(True)
            single = SamplesLoss(loss, **kw)(xs, ys)
            gxs, gys = torch.autograd.grad(single, [xs, ys])
            eng = ColumnShardedEngine()
            xg, yg = xh.clone().requires_grad_(True), yh.clone().requires_grad_(True)
            val = eng.attach(SamplesLoss(loss, **kw))(xg, yg)
            gx, gy = torch.autograd.grad(val, [xg, yg])
            vals = [None] * world
            dist.all_gather_object(vals, (val.item(), gx.abs().sum().item()))
            out[tag] = dict(val=val.item(), single=single.item(),
                            gx=(gx - gxs).abs().max().item() / gxs.abs().max().item(),
                            gy=(gy - gys).abs().max().item() / gys.abs().max().item(),
                            replicated=all(v == vals[0] for v in vals), collectives=eng.collectives, tol=2e-5)
        # two-scale Sinkhorn (BASELINE configs[3]): coarse phase replicated, block-sparse fine phase sharded by column tiles
        gm = torch.Generator().manual_seed(3)
        xm = torch.rand(40000, 3, generator=gm).to(dev)
        ym = (torch.rand(36000, 3, generator=gm) * 0.8 + 0.1).to(dev)
        for tag, kw in (("ms_trunc", dict(p=2, blur=0.02, truncate=5)), ("ms_exact", dict(p=2, blur=0.03, truncate=None))):
            xs = xm.clone().requires_grad_(True)
            single = SamplesLoss("sinkhorn", backend="multiscale", cluster_scale=0.12, **kw)(xs, ym)
            (gxs,) = torch.autograd.grad(single, xs)
            eng = ColumnShardedEngine()
            xg = xm.clone().requires_grad_(True)
            val = eng.attach(SamplesLoss("sinkhorn", backend="multiscale", cluster_scale=0.12, **kw))(xg, ym)
            (gx,) = torch.autograd.grad(val, xg)
            vals = [None] * world
            dist.all_gather_object(vals, (val.item(), gx.abs().sum().item()))
            out[tag] = dict(val=val.item(), single=single.item(),
                            gx=(gx - gxs).abs().max().item() / gxs.abs().max().item(), gy=0.0,
                            replicated=all(v == vals[0] for v in vals), collectives=eng.collectives, tol=2e-5)
        if rank == 0:
            results.put(out)
    except Exception as exc:
        import traceback

        results.put({"error": f"rank {rank}: {exc!r}\n{traceback.format_exc()}"})
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_column_sharded_nccl():
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    out = results.get(timeout=600)
    assert "error" not in out, out.get("error")
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for tag, r in out.items():
        assert r["replicated"], r
        # (multiscale: centroids come from float atomics, so two runs differ by a few ulps of the potentials)
        assert abs(r["val"] - r["single"]) <= r.get("tol", 2e-6) * abs(r["single"]) + 1e-9, r
        # (truncated two-scale runs may keep slightly different tile sets: borderline cluster pairs flip with the
        #  last bits of the atomically accumulated centroids; their weight is ~exp(-truncate))
        gtol = 1e-3 if tag in ("ms_trunc", "d16_gaussian") else 1e-4
        assert r["gx"] < gtol and r["gy"] < gtol, r
        assert r["collectives"] > 0
