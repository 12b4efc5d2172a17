"""BASELINE.json configs[3] (multiscale Sinkhorn, one GPU's view) and configs[4] (unbalanced Sinkhorn on a
256^3 volume) — SURVEY.md section 8(d) input recipes.

    python tools/bench_configs.py multiscale 1000000 [10000000]
    python tools/bench_configs.py grid 128 256
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from geomloss_b200 import SamplesLoss, sinkhorn_divergence, ops  # noqa: E402
from geomloss_b200.sinkhorn_images import softmin_grid  # noqa: E402

dev = "cuda:0"  # (the sharded benchmark uses cuda:LOCAL_RANK)


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        l0 = ops.launches()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        nl = ops.launches() - l0
    return best, out, nl


def multiscale(N):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 3, generator=g).to(dev)
    y = torch.rand(N, 3, generator=g).to(dev)
    for blur in (0.05, 0.01):
        L = SamplesLoss("sinkhorn", p=2, blur=blur, scaling=0.5, truncate=5, backend="multiscale")
        t, v, nl = timed(lambda: L(x, y), reps=2)
        print(json.dumps({"config": "multiscale sinkhorn", "N": N, "D": 3, "blur": blur, "scaling": 0.5, "truncate": 5,
                          "s_fwd": round(t, 4), "launches": nl, "value": float(v),
                          "mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
        xg = x.clone().requires_grad_(True)

        def fb():
            val = L(xg, y)
            torch.autograd.grad(val, xg)
            return val

        t, v, nl = timed(fb, reps=1)
        print(json.dumps({"config": "multiscale sinkhorn", "N": N, "blur": blur, "s_fwd_bwd": round(t, 4)}), flush=True)


def blobs(n, seed, mass):
    g = torch.Generator().manual_seed(seed)
    ax = torch.linspace(0, 1, n)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    f = torch.full((n, n, n), 1e-3)
    for _ in range(3):
        c = 0.25 + 0.5 * torch.rand(3, generator=g)
        s = 0.05 + 0.1 * float(torch.rand(1, generator=g))
        f = f + torch.exp(-((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) / (2 * s * s))
    return (f * (mass / f.sum()))[None, None].to(dev)


def grid(n):
    a, b = blobs(n, 0, 1.0), blobs(n, 1, 1.3)
    # the separable softmin alone: one call = 3 axis passes, n^4 pairs each, 8 n^3 bytes of HBM traffic each
    h = torch.log(a)
    f = torch.zeros_like(a)
    eps = (1.0 / n) ** 2
    t, _, _ = timed(lambda: softmin_grid(eps, 2, h, f, 1.0 / eps), reps=5)
    print(json.dumps({"op": "softmin_grid", "n": n, "ms": round(t * 1e3, 3), "pairs_per_s": 3 * n**4 / t,
                      "hbm_GBps_algorithmic": 3 * 8 * n**3 / t / 1e9}), flush=True)
    for reach in (0.3, None):
        t, v, nl = timed(lambda: sinkhorn_divergence(a, b, p=2, blur=1.0 / n, reach=reach, scaling=0.5), reps=2)
        print(json.dumps({"config": "sinkhorn_images.sinkhorn_divergence", "grid": [n, n, n], "reach": reach,
                          "blur": 1.0 / n, "scaling": 0.5, "s": round(t, 4), "launches": nl, "value": float(v[0])}),
              flush=True)


def multiscale_sharded(N):
    """torchrun --nproc-per-node W tools/bench_configs.py multiscale_sharded N   (BASELINE configs[3])"""
    import os

    import torch.distributed as dist

    from geomloss_b200.distributed import ColumnShardedEngine

    rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=d)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 3, generator=g).to(d)
    y = torch.rand(N, 3, generator=g).to(d)
    for blur in (0.05, 0.01):
        eng = ColumnShardedEngine()
        L = eng.attach(SamplesLoss("sinkhorn", p=2, blur=blur, scaling=0.5, truncate=5, backend="multiscale"))
        for rep in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            v = L(x, y)
            torch.cuda.synchronize()
            dist.barrier()
            t = time.perf_counter() - t0
        tt = torch.tensor([t], device=d)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"config": "multiscale sinkhorn, column-sharded", "world": dist.get_world_size(), "N": N,
                              "blur": blur, "scaling": 0.5, "truncate": 5, "s_fwd": round(float(tt), 4),
                              "value": float(v), "collectives": eng.collectives}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    for arg in sys.argv[2:]:
        {"multiscale": multiscale, "grid": grid, "multiscale_sharded": multiscale_sharded}[what](int(arg))
