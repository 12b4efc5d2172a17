"""GPU debugging aid: small-path vs tiled-path vs fp64 torch for p = 1 (clamped / unclamped)."""
import sys

import torch

sys.path.insert(0, ".")
from geomloss_b200 import SamplesLoss, sinkhorn_small  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
n, m, d = 700, 800, 3
x, y = torch.rand(n, d, generator=g).to(DEV), (torch.rand(m, d, generator=g) * 0.9 + 0.1).to(DEV)
a = torch.rand(n, generator=g).to(DEV) + 0.1
b = torch.rand(m, generator=g).to(DEV) + 0.1
a, b = a / a.sum(), b / b.sum()


def ref64(backend, p, blur, reach=None):
    """fp64 torch restatement on the device of the dense loss (keops = unclamped sqrt when backend != tensorized)."""
    sys.path.insert(0, ".")
    from oracle import geomloss_oracle as O

    xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
    if backend == "tensorized":
        val = O.samples_loss(a.double(), xr, b.double(), yr, loss="sinkhorn", p=p, blur=blur, reach=reach)
    else:
        val = O.sinkhorn_online(a.double()[None], xr[None], b.double()[None], yr[None], p=p, blur=blur, reach=reach)[0]
    gx, gy = torch.autograd.grad(val, [xr, yr])
    return val.item(), gx, gy


for backend in ("tensorized", "online"):
    for p, blur in ((1, 0.05), (2, 0.05)):
        rv, rgx, rgy = ref64(backend, p, blur)
        for path, small_max in (("small", 6000), ("tiled", 0)):
            sinkhorn_small.SMALL_MAX = small_max
            xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            v = SamplesLoss("sinkhorn", p=p, blur=blur, backend=backend)(a, xg, b, yg)
            gx, gy = torch.autograd.grad(v, [xg, yg])
            F, G = SamplesLoss("sinkhorn", p=p, blur=blur, backend=backend, potentials=True)(a, x, b, y)
            print(f"{backend:10s} p={p} {path:5s} value rel {abs(v.item() - rv) / abs(rv):.2e}  "
                  f"gx relmax {(gx.double() - rgx).abs().max().item() / rgx.abs().max().item():.2e}  "
                  f"gy relmax {(gy.double() - rgy).abs().max().item() / rgy.abs().max().item():.2e}  "
                  f"F[0:3] {F.flatten()[:3].tolist()}")
sinkhorn_small.SMALL_MAX = 6000
