"""Golden fixtures in dimensions 16 and 64 (the tensor-core kernels) from the REAL reference's tensorized backend.

    python tests/golden/make_golden_highdim.py

Imports geomloss from /root/reference/src (commit 00e493f); fp32 and fp64 runs of
SamplesLoss("gaussian" | "sinkhorn", backend="tensorized") with autograd gradients.  The blur=.05, D=64 case is the
regime of BASELINE configs[2]: every off-diagonal kernel value underflows and the loss is 1/2 sum a_i^2 + 1/2 sum b_j^2,
carried by the diagonal of K_xx and K_yy.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("GEOMLOSS_REFERENCE", "/root/reference/src"))
from geomloss import SamplesLoss  # noqa: E402

torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy()


CASES = [
    ("hd_gaussian_d16_blur2", 41, 700, 600, 16, dict(loss="gaussian", blur=2.0)),
    ("hd_gaussian_d64_blur2", 42, 700, 600, 64, dict(loss="gaussian", blur=2.0)),
    ("hd_gaussian_d64_blur07", 43, 600, 650, 64, dict(loss="gaussian", blur=0.7)),
    ("hd_gaussian_d64_blur005", 44, 600, 650, 64, dict(loss="gaussian", blur=0.05)),
    ("hd_gaussian_d33_blur1", 45, 500, 450, 33, dict(loss="gaussian", blur=1.0)),
    ("hd_sinkhorn_d16", 46, 500, 550, 16, dict(loss="sinkhorn", p=2, blur=0.5, scaling=0.6)),
    ("hd_sinkhorn_d64", 47, 450, 500, 64, dict(loss="sinkhorn", p=2, blur=1.0, scaling=0.6)),
    ("hd_sinkhorn_d64_reach", 48, 450, 400, 64, dict(loss="sinkhorn", p=2, blur=1.0, reach=3.0, scaling=0.7)),
]


def main():
    for name, seed, n, m, d, kw in CASES:
        g = torch.Generator().manual_seed(seed)
        # fp32-representable inputs: the fp64 run of the reference sees exactly the same numbers (x.double())
        x = torch.rand(n, d, generator=g)
        y = torch.rand(m, d, generator=g) * 0.9 + 0.15
        a = torch.rand(n, generator=g) + 0.1
        b = torch.rand(m, generator=g) + 0.1
        a, b = a / a.sum(), b / b.sum()
        arrays = dict(a=npy(a), x=npy(x), b=npy(b), y=npy(y))
        for k, v in kw.items():
            arrays["kw_" + k] = np.array(v)
        for dtype, suf in ((torch.float32, ""), (torch.float64, "_f64")):
            leaves = [t.to(dtype).requires_grad_(True) for t in (a, x, b, y)]
            val = SamplesLoss(backend="tensorized", **kw)(*leaves)
            ga, gx, gb, gy = torch.autograd.grad(val, leaves)
            arrays.update({"value" + suf: npy(val), "grad_a" + suf: npy(ga), "grad_b" + suf: npy(gb)})
            if suf:  # point gradients: the fp64 run only (the fp32 run's are within its own rounding noise of it)
                arrays.update({"grad_x" + suf: npy(gx).astype(np.float32), "grad_y" + suf: npy(gy).astype(np.float32)})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB value fp32 {arrays['value']} fp64 {arrays['value_f64']}")


if __name__ == "__main__":
    main()
