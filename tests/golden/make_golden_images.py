"""Golden fixtures for the reference's grid paths: ``softmin_grid``, ``sinkhorn_images.sinkhorn_divergence`` and
``ImagesBarycenter``.

    python tests/golden/make_golden_images.py

The UNMODIFIED reference under /root/reference/src (commit 00e493f) runs on ``tests/golden/pykeops_shim`` (dense
torch stand-in for the LazyTensor reductions of _legacy/utils.py:247-259).  fp32 AND fp64 runs are stored.

Reference entry points exercised:
  softmin_grid             _legacy/utils.py:190-279
  sinkhorn_divergence      _legacy/sinkhorn_images.py:26-202 (pyramid/upsample/log_dens utils.py:88-108)
  ImagesBarycenter         _legacy/wasserstein_barycenter_images.py:6-93
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GEOMLOSS_REFERENCE", "/root/reference/src")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "pykeops_shim"))
import pykeops  # noqa: E402

assert pykeops.__version__.startswith("shim")
from geomloss import ImagesBarycenter, sinkhorn_divergence  # noqa: E402
from geomloss._legacy import utils as ref_utils  # noqa: E402

assert ref_utils.keops_available
torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy()


def blobs(seed, shape, floor=1e-3, zero_corner=False):
    """Non-negative fields: a few Gaussian bumps + a floor (BASELINE configs[4] style), normalised per (b, c)."""
    g = torch.Generator().manual_seed(seed)
    B, C, n = shape[0], shape[1], shape[-1]
    dim = len(shape) - 2
    ax = (torch.arange(n, dtype=torch.float64) + 0.5) / n
    grids = torch.meshgrid(*([ax] * dim), indexing="ij")
    out = torch.zeros(shape, dtype=torch.float64)
    for bi in range(B):
        for ci in range(C):
            f = torch.full(shape[2:], floor, dtype=torch.float64)
            for _ in range(3):
                c = torch.rand(dim, generator=g, dtype=torch.float64) * 0.6 + 0.2
                s = 0.05 + 0.1 * torch.rand(1, generator=g, dtype=torch.float64)
                f = f + torch.exp(-sum((gr - c[k]) ** 2 for k, gr in enumerate(grids)) / (2 * s**2))
            if zero_corner:
                f[(slice(0, n // 4),) * dim] = 0.0  # exercises the -10000 floor of log_dens
            out[bi, ci] = f / f.sum()
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def run_divergence(a, b, dtype, potentials, kw):
    a, b = a.to(dtype), b.to(dtype)
    if potentials:
        F, G = sinkhorn_divergence(a, b, potentials=True, **kw)
        return dict(pot_f=npy(F), pot_g=npy(G))
    ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    val = sinkhorn_divergence(ag, bg, **kw)
    ga, gb = torch.autograd.grad(val.sum(), [ag, bg])
    return dict(value=npy(val), grad_a=npy(ga), grad_b=npy(gb))


DIVERGENCE_CASES = [
    # name, seed, shape, mass of b, kwargs, potentials, zero corner
    ("img_div_2d_32", 1, (2, 1, 32, 32), 1.0, dict(p=2), False, False),
    ("img_div_2d_64_blur", 2, (1, 2, 64, 64), 1.0, dict(p=2, blur=0.03, scaling=0.7), False, False),
    ("img_div_2d_32_p1", 3, (1, 1, 32, 32), 1.0, dict(p=1, blur=0.05), False, False),
    ("img_div_2d_32_reach", 4, (2, 2, 32, 32), 1.3, dict(p=2, reach=0.3), False, False),
    ("img_div_2d_32_nodebias", 5, (1, 1, 32, 32), 1.0, dict(p=2, debias=False, blur=0.05), False, False),
    ("img_div_2d_32_potentials", 6, (1, 2, 32, 32), 1.0, dict(p=2, blur=0.04), True, False),
    ("img_div_2d_16_zeros", 7, (1, 1, 16, 16), 1.0, dict(p=2), False, True),
    ("img_div_3d_16", 8, (1, 1, 16, 16, 16), 1.0, dict(p=2), False, False),
    ("img_div_3d_32_reach", 9, (1, 1, 32, 32, 32), 1.3, dict(p=2, reach=0.3), False, False),
    ("img_div_3d_16_p1_potentials", 10, (1, 2, 16, 16, 16), 1.2, dict(p=1, reach=0.5, blur=0.08), True, False),
    ("img_div_3d_16_scaling", 11, (2, 1, 16, 16, 16), 1.0, dict(p=2, scaling=0.8, blur=0.1), False, False),
]


def main():
    # ---- operator level: softmin_grid(eps, p, h) ----
    g = torch.Generator().manual_seed(0)
    ops = {}
    for tag, shape, p, eps in [("2d_p2", (2, 2, 32, 32), 2, 0.01), ("2d_p1", (1, 2, 16, 16), 1, 0.05),
                               ("3d_p2", (1, 2, 16, 16, 16), 2, 0.004), ("3d_p1", (1, 1, 8, 8, 8), 1, 0.2),
                               ("2d_p2_64_sharp", (1, 1, 64, 64), 2, (1 / 64) ** 2)]:
        h = torch.randn(*shape, generator=g, dtype=torch.float64) * 3.0
        ops[f"{tag}_h"] = npy(h)
        ops[f"{tag}_p"], ops[f"{tag}_eps"] = np.array(p), np.array(eps)
        ops[f"{tag}_out"] = npy(ref_utils.softmin_grid(eps, p, h.float()))
        ops[f"{tag}_out_f64"] = npy(ref_utils.softmin_grid(eps, p, h))
    save("img_softmin_grid_operator", **ops)

    # ---- sinkhorn_images.sinkhorn_divergence ----
    for name, seed, shape, mass_b, kw, pot, zc in DIVERGENCE_CASES:
        a, b = blobs(seed, shape, zero_corner=zc), blobs(seed + 100, shape) * mass_b
        arrays = dict(a=npy(a.float()), b=npy(b.float()), a_f64=npy(a), b_f64=npy(b), potentials=int(pot))
        for k, v in kw.items():
            arrays["kw_" + k] = np.array(v)
        r32, r64 = run_divergence(a, b, torch.float32, pot, kw), run_divergence(a, b, torch.float64, pot, kw)
        arrays.update(r32)
        arrays.update({k + "_f64": v for k, v in r64.items()})
        save(name, **arrays)

    # ---- ImagesBarycenter ----
    for name, seed, shape, kw in [("img_bary_16", 21, (2, 3, 16, 16), dict()),
                                  ("img_bary_32_blur", 22, (1, 2, 32, 32), dict(blur=0.05, scaling_N=6)),
                                  ("img_bary_16_p1", 23, (1, 3, 16, 16), dict(p=1, blur=0.1, scaling_N=5,
                                                                            backward_iterations=3)),
                                  ("img_bary_16_fullgrad", 24, (1, 2, 16, 16), dict(scaling_N=4,
                                                                                  backward_iterations=0))]:
        meas = blobs(seed, shape)
        gw = torch.Generator().manual_seed(seed)
        w = torch.rand(shape[0], shape[1], generator=gw, dtype=torch.float64) + 0.2
        w = w / w.sum(1, keepdim=True)
        probe = torch.randn(shape[0], 1, shape[2], shape[3], generator=gw, dtype=torch.float64)
        arrays = dict(measures_f64=npy(meas), weights_f64=npy(w), probe_f64=npy(probe))
        for k, v in kw.items():
            arrays["kw_" + k] = np.array(v)
        for dtype, suf in ((torch.float32, ""), (torch.float64, "_f64")):
            m, ww = meas.to(dtype).requires_grad_(True), w.to(dtype).requires_grad_(True)
            bar = ImagesBarycenter(m, ww, **kw)
            gm, gww = torch.autograd.grad((bar * probe.to(dtype)).sum(), [m, ww], allow_unused=True)
            arrays.update({"bar" + suf: npy(bar), "grad_weights" + suf: npy(gww)})
            # with backward_iterations > 0 the pyramid of the measures is built without a graph (:43-51):
            # the reference carries NO gradient to `measures` then
            arrays["measures_have_grad"] = int(gm is not None)
            if gm is not None:
                arrays["grad_measures" + suf] = npy(gm)
        save(name, **arrays)


if __name__ == "__main__":
    main()
