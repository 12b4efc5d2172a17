"""Golden fixtures for the reference's pykeops-backed point-cloud paths: ``backend="multiscale"`` (Sinkhorn and
kernel MMD) and ``backend="online"``.

    python tests/golden/make_golden_multiscale.py

The UNMODIFIED reference under /root/reference/src (commit 00e493f) is imported with ``tests/golden/pykeops_shim``
first on ``sys.path``: a dense pure-torch stand-in for the pykeops calls the reference makes (see the shim's
docstring for the call-site list).  Every case is run in fp32 AND fp64; inputs, hyper-parameters and outputs
(value, autograd gradients w.r.t. weights and points, or the potentials) are stored as ``ms_*.npz``.

Reference entry points exercised:
  sinkhorn_multiscale      _legacy/sinkhorn_samples.py:547-681  (clusterize :453-490, kernel_truncation :493-530,
                           extrapolate_samples :533-544, un-permutation :675-679, jump branch of sinkhorn_loop
                           _legacy/sinkhorn_divergence.py:519-606)
  kernel_multiscale        _legacy/kernel_samples.py:177-271
  sinkhorn_online          _legacy/sinkhorn_samples.py:349-424 (generic_logsumexp and batched LazyTensor forms)
  kernel_online            _legacy/kernel_samples.py:160
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GEOMLOSS_REFERENCE", "/root/reference/src")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "pykeops_shim"))
import pykeops  # noqa: E402

assert pykeops.__version__.startswith("shim"), "the dense pykeops shim must shadow any real pykeops"
from geomloss import SamplesLoss  # noqa: E402
from geomloss._legacy import sinkhorn_samples as ref_ss  # noqa: E402

assert ref_ss.keops_available
torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy()


def clouds(seed, n, m, d, weights="uniform", batch=None):
    g = torch.Generator().manual_seed(seed)
    sx = (n, d) if batch is None else (batch, n, d)
    sy = (m, d) if batch is None else (batch, m, d)
    x = torch.rand(*sx, generator=g, dtype=torch.float64)
    y = torch.rand(*sy, generator=g, dtype=torch.float64) * 0.9 + 0.15
    if weights == "uniform":
        a = torch.ones(sx[:-1], dtype=torch.float64) / n
        b = torch.ones(sy[:-1], dtype=torch.float64) / m
    else:
        a = torch.rand(sx[:-1], generator=g, dtype=torch.float64) + 0.1
        b = torch.rand(sy[:-1], generator=g, dtype=torch.float64) + 0.1
        a, b = a / a.sum(-1, keepdim=True), b / b.sum(-1, keepdim=True)
        if weights == "unbalanced":
            b = b * 1.3
    return a, x, b, y


def coarse_labels(x, cell):
    """User-supplied cluster labels (the ``labels_x=`` argument form): a coarse voxel grid, compact ids."""
    ij = torch.floor(x / cell).long()
    key = ij[:, 0]
    for k in range(1, x.shape[1]):
        key = key * 64 + ij[:, k]
    return torch.unique(key, return_inverse=True)[1].int()


def run(a, x, b, y, dtype, labels=None, potentials=False, **kw):
    a, x, b, y = (t.to(dtype) for t in (a, x, b, y))
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        L = SamplesLoss(potentials=potentials, **kw)
        args = (a, x, b, y) if labels is None else (labels[0], a, x, labels[1], b, y)
        if potentials:
            F, G = L(*args)
            out.update(pot_f=npy(F), pot_g=npy(G))
        else:
            ag, xg, bg, yg = (t.clone().requires_grad_(True) for t in (a, x, b, y))
            args = (ag, xg, bg, yg) if labels is None else (labels[0], ag, xg, labels[1], bg, yg)
            val = L(*args)
            ga, gx, gb, gy = torch.autograd.grad(val.sum(), [ag, xg, bg, yg])
            out.update(value=npy(val), grad_a=npy(ga), grad_x=npy(gx), grad_b=npy(gb), grad_y=npy(gy))
    return out


def save(name, a, x, b, y, kw, labels=None, potentials=False):
    arrays = dict(a=npy(a.float()), x=npy(x.float()), b=npy(b.float()), y=npy(y.float()))
    arrays.update(a_f64=npy(a), x_f64=npy(x), b_f64=npy(b), y_f64=npy(y))
    if labels is not None:
        arrays.update(labels_x=npy(labels[0]), labels_y=npy(labels[1]))
    arrays["potentials"] = int(potentials)
    for k, v in kw.items():
        arrays["kw_" + k] = np.array("None" if v is None else v)
    r32 = run(a, x, b, y, torch.float32, labels=labels, potentials=potentials, **kw)
    r64 = run(a, x, b, y, torch.float64, labels=labels, potentials=potentials, **kw)
    arrays.update(r32)
    arrays.update({k + "_f64": v for k, v in r64.items()})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    key = "pot_f" if potentials else "value"
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  {key} fp32 {np.ravel(r32[key])[:1]} fp64 {np.ravel(r64[key])[:1]}")


SINKHORN_CASES = [
    # (name suffix, seed, N, M, D, weights, SamplesLoss kwargs, potentials, labels cell or None)
    ("d3_default", 1, 1400, 1600, 3, "uniform", dict(p=2, blur=0.05, truncate=5), False, None),
    ("d3_trunc1", 2, 1400, 1500, 3, "random", dict(p=2, blur=0.05, truncate=1, cluster_scale=0.15), False, None),
    ("d3_exact", 3, 1200, 1300, 3, "random", dict(p=2, blur=0.05, truncate=None, cluster_scale=0.12), False, None),
    ("d2_trunc_half", 4, 1500, 1400, 2, "random", dict(p=2, blur=0.03, truncate=0.5, cluster_scale=0.08), False, None),
    ("d1", 5, 900, 1100, 1, "random", dict(p=2, blur=0.02, truncate=2, cluster_scale=0.05), False, None),
    ("d3_p1", 6, 1200, 1300, 3, "random", dict(p=1, blur=0.05, truncate=2, cluster_scale=0.15), False, None),
    ("d3_reach", 7, 1300, 1200, 3, "unbalanced", dict(p=2, blur=0.05, reach=0.3, truncate=3, cluster_scale=0.15), False, None),
    ("d3_nodebias", 8, 1300, 1200, 3, "random", dict(p=2, blur=0.05, truncate=3, cluster_scale=0.15, debias=False), False, None),
    ("d3_potentials", 9, 1300, 1200, 3, "random", dict(p=2, blur=0.05, truncate=3, cluster_scale=0.15), True, None),
    ("d2_potentials_reach_nodebias", 10, 1100, 1200, 2, "unbalanced",
     dict(p=2, blur=0.05, reach=0.5, truncate=3, cluster_scale=0.1, debias=False), True, None),
    ("d3_lastjump", 11, 1100, 1000, 3, "random", dict(p=2, blur=0.1, truncate=5, cluster_scale=0.05), False, None),
    ("d3_lastjump_potentials", 12, 1000, 1100, 3, "random", dict(p=2, blur=0.1, truncate=5, cluster_scale=0.05), True, None),
    ("d3_labels", 13, 1200, 1300, 3, "random", dict(p=2, blur=0.05, truncate=2, cluster_scale=0.2), False, 0.2),
    ("d3_fine_schedule", 14, 1000, 1000, 3, "uniform", dict(p=2, blur=0.02, scaling=0.7, truncate=1, cluster_scale=0.12), False, None),
    ("d3_reach_p1", 15, 1000, 1100, 3, "unbalanced", dict(p=1, blur=0.05, reach=0.4, truncate=3, cluster_scale=0.15), False, None),
]

KERNEL_CASES = [
    ("gaussian_bench", 21, 1500, 1600, 3, "random", dict(loss="gaussian", blur=0.1, truncate=3), False),
    ("gaussian_d2", 22, 1400, 1300, 2, "random", dict(loss="gaussian", blur=0.05, truncate=2), False),
    ("gaussian_cluster_scale", 23, 1200, 1300, 3, "uniform", dict(loss="gaussian", blur=0.1, truncate=1, cluster_scale=1.0), False),
    ("gaussian_diameter", 24, 1200, 1300, 3, "random", dict(loss="gaussian", blur=0.1, truncate=2, diameter=3.0), False),
    ("laplacian", 25, 1300, 1200, 3, "random", dict(loss="laplacian", blur=0.1, truncate=5), False),
    ("energy", 26, 1000, 1100, 3, "random", dict(loss="energy", truncate=5), False),
    ("gaussian_notrunc", 27, 1000, 1100, 3, "random", dict(loss="gaussian", blur=0.1, truncate=None), False),
    ("gaussian_potentials", 28, 1200, 1300, 3, "random", dict(loss="gaussian", blur=0.1, truncate=2), True),
]


def main():
    for suffix, seed, n, m, d, w, kw, pot, cell in SINKHORN_CASES:
        a, x, b, y = clouds(seed, n, m, d, w)
        labels = None if cell is None else (coarse_labels(x, cell), coarse_labels(y, cell))
        save("ms_sinkhorn_" + suffix, a, x, b, y, dict(loss="sinkhorn", backend="multiscale", **kw), labels=labels,
             potentials=pot)
    for suffix, seed, n, m, d, w, kw, pot in KERNEL_CASES:
        a, x, b, y = clouds(seed, n, m, d, w)
        save("ms_kernel_" + suffix, a, x, b, y, dict(backend="multiscale", **kw), potentials=pot)
    # backend="online": unbatched (generic_logsumexp) and batched (LazyTensor) forms
    a, x, b, y = clouds(31, 700, 800, 3, "random")
    save("online_sinkhorn_p2", a, x, b, y, dict(loss="sinkhorn", backend="online", p=2, blur=0.05))
    save("online_sinkhorn_p1", a, x, b, y, dict(loss="sinkhorn", backend="online", p=1, blur=0.05, reach=0.5))
    a, x, b, y = clouds(32, 300, 350, 2, "random", batch=3)
    save("online_sinkhorn_batched", a, x, b, y, dict(loss="sinkhorn", backend="online", p=2, blur=0.05))
    save("online_sinkhorn_batched_p1", a, x, b, y, dict(loss="sinkhorn", backend="online", p=1, blur=0.1))
    save("online_gaussian_batched", a, x, b, y, dict(loss="gaussian", backend="online", blur=0.2))
    save("online_laplacian_batched", a, x, b, y, dict(loss="laplacian", backend="online", blur=0.2))
    save("online_energy", a[0], x[0], b[0], y[0], dict(loss="energy", backend="online"))


if __name__ == "__main__":
    main()
