"""The CUDA engine against the LIVE, unmodified reference on fresh random inputs (not the committed fixtures).

``oracle/_ref/`` holds jeanfeydy/geomloss @ 00e493f, installed by ``oracle/make_ref.sh`` in the build container; it is
git-ignored but travels to the GPU box with the snapshot.  The reference runs on the CPU here: its tensorized backend
as is, its pykeops-backed backends (online / multiscale / images) on ``tests/golden/pykeops_shim``.  Skipped when
``oracle/_ref`` is absent.  Tolerances as in tests/test_gpu_reference_goldens.py, against the reference's fp64 run.
"""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(os.path.join(REF_DIR, "geomloss")):
        pytest.skip("oracle/_ref is not installed (run oracle/make_ref.sh in the build container)")
    shim = os.path.join(GOLDEN, "pykeops_shim")
    sys.path.insert(0, REF_DIR)
    sys.path.insert(0, shim)
    for mod in [m for m in sys.modules if m == "geomloss" or m.startswith("geomloss.") or m.startswith("pykeops")]:
        del sys.modules[mod]
    import geomloss

    assert os.path.realpath(geomloss.__file__).startswith(os.path.realpath(REF_DIR))
    yield geomloss
    sys.path.remove(REF_DIR)
    sys.path.remove(shim)
    for mod in [m for m in sys.modules if m == "geomloss" or m.startswith("geomloss.") or m.startswith("pykeops")]:
        del sys.modules[mod]


def _clouds(seed, n, m, d):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = torch.rand(m, d, generator=g, dtype=torch.float64) * 0.8 + 0.2
    a = torch.rand(n, generator=g, dtype=torch.float64) + 0.2
    b = torch.rand(m, generator=g, dtype=torch.float64) + 0.2
    return a / a.sum(), x, b / b.sum(), y


def test_installed_reference_reproduces_a_committed_fixture(ref):
    """oracle/_ref is the code that generated tests/golden: bit-for-bit on cfg1 (fp32 run)."""
    with np.load(os.path.join(GOLDEN, "cfg1_sinkhorn_n1000.npz")) as g:
        x, y, want = torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), float(g["value_f32"])
    got = ref.SamplesLoss("sinkhorn", p=2, blur=0.05, backend="tensorized")(x, y).item()
    assert got == pytest.approx(want, rel=1e-6)


CASES = [
    ("tensorized", dict(loss="sinkhorn", p=2, blur=0.05), 3),
    ("tensorized", dict(loss="sinkhorn", p=1, blur=0.05, reach=0.4), 2),
    ("online", dict(loss="sinkhorn", p=1, blur=0.05), 3),
    ("multiscale", dict(loss="sinkhorn", p=2, blur=0.03, truncate=1, cluster_scale=0.1), 3),
    ("multiscale", dict(loss="sinkhorn", p=2, blur=0.05, reach=0.5, truncate=3), 2),
    ("tensorized", dict(loss="gaussian", blur=0.2), 3),
    ("online", dict(loss="laplacian", blur=0.2), 3),
    ("multiscale", dict(loss="gaussian", blur=0.1, truncate=2), 3),
    ("tensorized", dict(loss="energy"), 5),
]


@pytest.mark.gpu
@pytest.mark.parametrize("backend,kw,d", CASES)
def test_samples_loss_vs_live_reference(ref, backend, kw, d, sinkhorn_path):
    from geomloss_b200 import SamplesLoss

    a, x, b, y = _clouds(sum(map(ord, backend + kw["loss"])) + d, 1700, 1500, d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        leaves = [t.clone().requires_grad_(True) for t in (a, x, b, y)]
        want = ref.SamplesLoss(backend=backend, **kw)(*leaves)
        want_g = torch.autograd.grad(want.sum(), leaves)
    dl = [t.float().to(DEV).requires_grad_(True) for t in (a, x, b, y)]
    got = SamplesLoss(backend=backend, **kw)(*dl)
    got_g = torch.autograd.grad(got.sum(), dl)
    assert tuple(got.shape) == tuple(want.shape)
    slack = 3e-7 * float((a.sum() ** 2 + b.sum() ** 2) / 2) if kw["loss"] != "sinkhorn" else 0.0
    assert abs(got.sum().item() - want.sum().item()) <= 1e-4 * abs(want.sum().item()) + slack
    for gg, wg in zip(got_g, want_g):
        assert (gg.cpu().double() - wg).abs().max().item() <= 5e-4 * wg.abs().max().item() + 1e-12


@pytest.mark.gpu
def test_image_divergence_vs_live_reference(ref):
    from geomloss_b200 import sinkhorn_divergence

    g = torch.Generator().manual_seed(11)
    a = torch.rand(1, 2, 32, 32, generator=g, dtype=torch.float64) ** 2
    b = torch.rand(1, 2, 32, 32, generator=g, dtype=torch.float64) ** 2
    a, b = a / a.sum(), 1.2 * b / b.sum()
    for kw in (dict(p=2, reach=0.4), dict(p=1, blur=0.06, scaling=0.6)):
        want = ref.sinkhorn_divergence(a, b, **kw)
        got = sinkhorn_divergence(a.float().to(DEV), b.float().to(DEV), **kw)
        np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(), rtol=1e-4)
