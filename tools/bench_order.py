"""Sensitivity of the softmin kernel to the ORDER of the points (the lazy running max re-bases whenever a chunk
holds a term 2^64 above the current reference): random order vs clouds sorted along x vs voxel(cluster)-sorted.

    python tools/bench_order.py [N] [eps ...]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from geomloss_b200 import ops  # noqa: E402
from geomloss_b200.multiscale import grid_labels  # noqa: E402

dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
EPS = [float(e) for e in sys.argv[2:]] or [1e-4, 2.5e-3]
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 3, generator=g).to(dev)
y = torch.rand(N, 3, generator=g).to(dev)
h = torch.full((N,), -float(torch.log(torch.tensor(float(N)))), device=dev)


def t_ms(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


orders = {
    "random": (x, y),
    "sorted_x": (x[x[:, 0].argsort()], y[y[:, 0].argsort()]),
    "voxel_sorted": (x[grid_labels(x, 0.08).argsort()], y[grid_labels(y, 0.08).argsort()]),
}
for eps in EPS:
    ref = None
    for name, (xs, ys) in orders.items():
        c = ops.default_center(xs, ys)
        ms = t_ms(lambda: ops.softmin_raw(eps, xs, ys, h, p=2, center=c))
        out = ops.softmin_raw(eps, xs, ys, h, p=2, center=c)[0]
        chk = float(out.double().sum())  # order-independent up to rounding
        ref = chk if ref is None else ref
        print(json.dumps({"order": name, "N": N, "eps": eps, "ms": round(ms, 2), "Tpairs_s": round(N * N / ms / 1e9, 3),
                          "sum_rel_diff": abs(chk - ref) / abs(ref)}), flush=True)
