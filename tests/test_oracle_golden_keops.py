"""Pin the oracle's pykeops-backed blocks (online / multiscale point clouds, grids, barycenters) against the
UNMODIFIED reference.

Fixtures: tests/golden/{ms_*,online_*,img_*}.npz, written by make_golden_multiscale.py / make_golden_images.py, which
run jeanfeydy/geomloss @ 00e493f from /root/reference/src on tests/golden/pykeops_shim (a dense torch stand-in for
the pykeops calls of the reference).  Every fixture holds the reference's fp32 and fp64 runs; the oracle is the same
dense arithmetic, so the fp64 comparison is at rounding level (1e-12) and pins every semantic choice: label order,
from_matrix blocks, un-permutation, the leaked eps of sinkhorn_multiscale, the unclamped KeOps Norm2, the
sorted-order potentials of kernel_multiscale.
"""
import numpy as np
import pytest
import torch

from conftest import golden_kwargs, golden_names, load_golden
from oracle import geomloss_oracle as O

T = torch.from_numpy
F64 = dict(rtol=1e-11, atol=1e-14)


def _clouds(g, dtype=torch.float64):
    return tuple(T(g[k + "_f64"]).to(dtype) for k in "axby")


def _run_points(g, dtype):
    kw = golden_kwargs(g)
    loss, backend = kw.pop("loss"), kw.pop("backend")
    pot = bool(g["potentials"])
    a, x, b, y = _clouds(g, dtype)
    if "labels_x" in g:
        kw.update(labels_x=T(g["labels_x"]), labels_y=T(g["labels_y"]))
    ag, xg, bg, yg = (t.clone().requires_grad_(not pot) for t in (a, x, b, y))
    batched = x.dim() == 3
    lift = (lambda t: t) if batched or backend == "multiscale" else (lambda t: t[None])
    args = tuple(lift(t) for t in (ag, xg, bg, yg))
    if loss == "sinkhorn" and backend == "multiscale":
        out = O.sinkhorn_multiscale_dense(*args, potentials=pot, **kw)
    elif backend == "multiscale":
        kw.pop("p", None)
        out = O.kernel_multiscale_dense(*args, loss, potentials=pot, **kw)
    elif loss == "sinkhorn":
        kw.pop("truncate", None)
        out = O.sinkhorn_online(*args, potentials=pot, **kw)
    else:
        kw.pop("truncate", None), kw.pop("p", None)
        out = O.mmd_online(*args, loss, potentials=pot, **kw)
    return out, (ag, xg, bg, yg), pot


@pytest.mark.parametrize("name", golden_names("ms_") + golden_names("online_"))
def test_point_cloud_keops_backends_fp64(name):
    g = load_golden(name)
    out, leaves, pot = _run_points(g, torch.float64)
    if pot:
        np.testing.assert_allclose(out[0].numpy().ravel(), g["pot_f_f64"].ravel(), **F64)
        np.testing.assert_allclose(out[1].numpy().ravel(), g["pot_g_f64"].ravel(), **F64)
        return
    np.testing.assert_allclose(out.detach().numpy().ravel(), g["value_f64"].ravel(), rtol=1e-11)
    grads = torch.autograd.grad(out.sum(), leaves)
    for got, key in zip(grads, ("grad_a", "grad_x", "grad_b", "grad_y")):
        ref = g[key + "_f64"]
        np.testing.assert_allclose(got.numpy().reshape(ref.shape), ref, rtol=1e-9, atol=1e-13 * max(np.abs(ref).max(), 1.0))


@pytest.mark.parametrize("name", ["ms_sinkhorn_d3_default", "ms_sinkhorn_d3_reach", "ms_sinkhorn_d3_p1",
                                  "ms_kernel_gaussian_bench", "online_sinkhorn_batched"])
def test_point_cloud_keops_backends_fp32(name):
    """fp32 run of the oracle vs the reference's fp32 run: same arithmetic, a few ulps."""
    g = load_golden(name)
    out, leaves, pot = _run_points(g, torch.float32)
    np.testing.assert_allclose(out.detach().numpy().ravel(), g["value"].ravel(), rtol=2e-6)
    grads = torch.autograd.grad(out.sum(), leaves)
    for got, key in zip(grads, ("grad_a", "grad_x", "grad_b", "grad_y")):
        ref = g[key]
        np.testing.assert_allclose(got.numpy().reshape(ref.shape), ref, atol=2e-5 * np.abs(ref).max())


def test_truncation_is_visible_in_the_fixtures():
    """The fixtures with a small ``truncate`` really exercise kernel truncation: the exact (truncate=None) value
    differs from the stored one by far more than the pinning tolerance."""
    g = load_golden("ms_sinkhorn_d2_trunc_half")
    kw = golden_kwargs(g)
    kw.pop("loss"), kw.pop("backend")
    a, x, b, y = _clouds(g)
    kw["truncate"] = None
    exact = O.sinkhorn_multiscale_dense(a, x, b, y, **kw).item()
    assert abs(exact - float(g["value_f64"])) > 1e-7 * abs(exact)


def test_kernel_multiscale_potentials_come_back_cluster_sorted():
    g = load_golden("ms_kernel_gaussian_potentials")
    a, x, b, y = _clouds(g)
    kw = golden_kwargs(g)
    F, G, lab_f, lab_g = O.kernel_multiscale_dense(a, x, b, y, "gaussian", blur=kw["blur"], truncate=kw["truncate"],
                                                   potentials=True)
    assert bool((lab_f[1:] >= lab_f[:-1]).all()) and bool((lab_g[1:] >= lab_g[:-1]).all())
    np.testing.assert_allclose(F.numpy(), g["pot_f_f64"], **F64)
    np.testing.assert_allclose(G.numpy(), g["pot_g_f64"], **F64)


# ---------------------------------------------------------------------------------------------- grids
def test_softmin_grid_operator():
    g = load_golden("img_softmin_grid_operator")
    for tag in ("2d_p2", "2d_p1", "3d_p2", "3d_p1", "2d_p2_64_sharp"):
        h, p, eps = T(g[tag + "_h"]), int(g[tag + "_p"]), float(g[tag + "_eps"])
        np.testing.assert_allclose(O.softmin_grid_dense(eps, p, h).numpy(), g[tag + "_out_f64"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(O.softmin_grid_dense(eps, p, h.float()).numpy(), g[tag + "_out"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", golden_names("img_div_"))
def test_image_divergence_fp64(name):
    g = load_golden(name)
    kw = golden_kwargs(g)
    pot = bool(g["potentials"])
    a, b = T(g["a_f64"]).requires_grad_(not pot), T(g["b_f64"]).requires_grad_(not pot)
    out = O.sinkhorn_images(a, b, potentials=pot, **kw)
    if pot:
        np.testing.assert_allclose(out[0].numpy(), g["pot_f_f64"], **F64)
        np.testing.assert_allclose(out[1].numpy(), g["pot_g_f64"], **F64)
        return
    np.testing.assert_allclose(out.detach().numpy(), g["value_f64"], rtol=1e-11)
    ga, gb = torch.autograd.grad(out.sum(), [a, b])
    np.testing.assert_allclose(ga.numpy(), g["grad_a_f64"], **F64)
    np.testing.assert_allclose(gb.numpy(), g["grad_b_f64"], **F64)


@pytest.mark.parametrize("name", golden_names("img_bary_"))
def test_images_barycenter_fp64(name):
    g = load_golden(name)
    kw = golden_kwargs(g)
    m, w = T(g["measures_f64"]).requires_grad_(True), T(g["weights_f64"]).requires_grad_(True)
    bar = O.images_barycenter(m, w, **kw)
    np.testing.assert_allclose(bar.detach().numpy(), g["bar_f64"], rtol=1e-10, atol=1e-15)
    gm, gw = torch.autograd.grad((bar * T(g["probe_f64"])).sum(), [m, w], allow_unused=True)
    np.testing.assert_allclose(gw.numpy(), g["grad_weights_f64"], rtol=1e-8, atol=1e-12)
    assert (gm is not None) == bool(g["measures_have_grad"])
    if gm is not None:
        np.testing.assert_allclose(gm.numpy(), g["grad_measures_f64"], rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------------ dimensions 16 .. 64 (tensorized reference)
@pytest.mark.parametrize("name", golden_names("hd_"))
def test_high_dimension_fixtures(name):
    g = load_golden(name)
    kw = golden_kwargs(g)
    leaves = [T(g[k]).double().requires_grad_(True) for k in "axby"]
    val = O.samples_loss(*leaves, **kw)
    np.testing.assert_allclose(val.item(), g["value_f64"], rtol=1e-11)
    ga, gx, gb, gy = torch.autograd.grad(val, leaves)
    np.testing.assert_allclose(ga.numpy(), g["grad_a_f64"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(gx.numpy(), g["grad_x_f64"], rtol=2e-6, atol=1e-7 * np.abs(g["grad_x_f64"]).max())
