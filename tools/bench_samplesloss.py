"""The reference's own benchmark protocol (examples/performances/plot_benchmarks_samplesloss_3D.py:31-111):
points on a sphere of diameter 1 in 3-D, random normalised weights, time = Loss(a, x, b, y) + backward,
one warm-up call then `loops` timed calls between synchronisations.

    python tools/bench_samplesloss.py [N ...]
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from geomloss_b200 import SamplesLoss  # noqa: E402

dev = "cuda:0"
CONFIGS = [("gaussian", dict(blur=0.1)), ("energy", dict()), ("sinkhorn", dict(p=2, blur=0.05, diameter=1.0)),
           ("sinkhorn", dict(p=2, blur=0.01, diameter=1.0))]


def sphere(n, g):
    x = torch.randn(n, 3, generator=g)
    x = x / (2 * x.norm(dim=1, keepdim=True))
    a = torch.rand(n, generator=g)
    return (a / a.sum()).to(dev), x.to(dev)


for N in [int(a) for a in sys.argv[1:]] or [1000, 10000, 100000]:
    g = torch.Generator().manual_seed(N)
    a, x = sphere(N, g)
    b, y = sphere(N, g)
    for loss, kw in CONFIGS:
        for backend in ("online", "multiscale") if loss == "sinkhorn" and N >= 100000 else ("online",):
            L = SamplesLoss(loss, backend=backend, **kw)
            xg = x.clone().requires_grad_(True)

            def call():
                val = L(a, xg, b, y)
                (gx,) = torch.autograd.grad(val, xg)
                return val

            call()
            torch.cuda.synchronize()
            loops = 10 if N <= 10000 else 3
            t0 = time.perf_counter()
            for _ in range(loops):
                v = call()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / loops
            print(json.dumps({"loss": loss, "backend": backend, **{k: v2 for k, v2 in kw.items()}, "N": N,
                              "s_per_call_fwd_bwd": round(dt, 6), "value": float(v)}))
