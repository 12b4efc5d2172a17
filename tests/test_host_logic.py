"""CPU tests of the host-side mirror of the reference interface (no kernel runs here):
schedule, scalar helpers, dual-to-value formulas, SamplesLoss argument handling and error behaviour."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from geomloss_b200 import SamplesLoss
from geomloss_b200 import sinkhorn as S
from oracle import geomloss_oracle as O


def test_epsilon_schedule_matches_reference_goldens():
    g = load_golden("epsilon_schedule")
    for i in range(4):
        p, diam, blur, scaling = g[f"args{i}"]
        got = np.array(S.epsilon_schedule(int(p), diam, blur, scaling))
        np.testing.assert_allclose(got, g[f"eps{i}"], rtol=1e-15)
    assert len(S.epsilon_schedule(2, 3**0.5, 0.01, 0.9)) == 51  # "~50 iters" of BASELINE cfg 2


def test_scalar_helpers_match_oracle():
    a = torch.tensor([0.0, 0.25, 0.75, -1.0])
    np.testing.assert_array_equal(S.log_weights(a).numpy(), O.log_weights(a.clone()).numpy())
    assert S.damping(0.3, None) == 1.0 and abs(S.damping(0.3, 0.2) - 1 / (1 + 0.3 / 0.2)) < 1e-15
    x, y = torch.rand(50, 3), torch.rand(40, 3) + 0.5
    assert abs(S.max_diameter(x, y) - O.max_diameter(x, y)) < 1e-7
    d, eps, lst, rho = S.scaling_parameters(x, y, 2, 0.05, 0.3, None, 0.5)
    d2, eps2, lst2, rho2 = O.scaling_parameters(x, y, 2, 0.05, 0.3, None, 0.5)
    assert (d, eps, rho) == (d2, eps2, rho2) and np.allclose(lst, lst2, rtol=1e-15)


@pytest.mark.parametrize("rho", [None, 0.09])
@pytest.mark.parametrize("debias", [True, False])
def test_sinkhorn_cost_matches_oracle(rho, debias):
    g = torch.Generator().manual_seed(3)
    n, m = 17, 23
    a, b = torch.rand(n, generator=g), torch.rand(m, generator=g)
    f_aa, f_ba = torch.rand(n, generator=g), torch.rand(n, generator=g)
    g_bb, g_ab = torch.rand(m, generator=g), torch.rand(m, generator=g)
    got = S.sinkhorn_cost(0.01, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias=debias)
    ref = O.sinkhorn_value(0.01, rho, a[None], b[None], f_aa[None], g_bb[None], g_ab[None], f_ba[None],
                           debias=debias)[0]
    assert abs(got.item() - ref.item()) < 1e-6
    F, G = S.sinkhorn_cost(0.01, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias=debias, potentials=True)
    Fr, Gr = O.sinkhorn_value(0.01, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias=debias, potentials=True)
    assert torch.equal(F, Fr) and torch.equal(G, Gr)


def test_constructor_keeps_reference_signature():
    L = SamplesLoss()
    assert (L.loss, L.p, L.blur, L.reach, L.diameter, L.scaling, L.truncate, L.cost, L.kernel, L.cluster_scale,
            L.debias, L.potentials, L.verbose, L.backend) == ("sinkhorn", 2, 0.05, None, None, 0.5, 5, None, None,
                                                              None, True, False, False, "auto")
    L = SamplesLoss("gaussian", blur=0.3, backend="online", potentials=True)
    assert L.loss == "gaussian" and L.blur == 0.3 and L.backend == "online" and L.potentials


def test_argument_forms_and_shape_errors():
    L = SamplesLoss()
    x, y = torch.rand(5, 3), torch.rand(7, 3)
    with pytest.raises(ValueError, match="two .* four .* or six"):
        L.process_args(x, y, x)
    l_x, a, xx, l_y, b, yy = L.process_args(x, y)
    assert l_x is None and l_y is None and torch.allclose(a, torch.full((5,), 0.2)) and b.shape == (7,)
    assert L.generate_weights(torch.rand(2, 4, 3)).shape == (2, 4)
    with pytest.raises(ValueError):
        L.generate_weights(torch.rand(4))
    # shape rules (samples_loss.py:337-474)
    with pytest.raises(ValueError, match="same last dimension"):
        L.check_shapes(None, a, x, None, b, torch.rand(7, 2))
    with pytest.raises(ValueError, match="same number of dimensions"):
        L.check_shapes(None, a, x, None, b, torch.rand(1, 7, 3))
    with pytest.raises(ValueError, match="compatible shapes"):
        L.check_shapes(None, torch.rand(4), x, None, b, y)
    with pytest.raises(ValueError, match=r"\(N,\) or \(N,1\)"):
        L.check_shapes(None, torch.rand(5, 2), x, None, torch.rand(7, 2), y)
    B, N, M, D, _, a2, _, b2 = L.check_shapes(None, a.view(-1, 1), x, None, b.view(-1, 1), y)
    assert (B, N, M, D) == (0, 5, 7, 3) and a2.shape == (5,) and b2.shape == (7,)
    xb, yb = torch.rand(2, 5, 3), torch.rand(2, 7, 3)
    B, N, M, D, _, a3, _, b3 = L.check_shapes(None, torch.rand(2, 5, 1), xb, None, torch.rand(2, 7, 1), yb)
    assert (B, N, M, D) == (2, 5, 7, 3) and a3.shape == (2, 5)
    with pytest.raises(ValueError, match="same batchsize"):
        L.check_shapes(None, torch.rand(2, 5), xb, None, torch.rand(2, 7), torch.rand(3, 7, 3))
    with pytest.raises(NotImplementedError):
        L.check_shapes(torch.zeros(2, 5), torch.rand(2, 5), xb, None, torch.rand(2, 7), yb)
    with pytest.raises(ValueError, match="labels 'l_x'"):
        L.check_shapes(torch.zeros(4), a, x, None, b, y)


def test_routing_errors_match_reference():
    x, y = torch.rand(5, 3), torch.rand(7, 3)
    a, b = torch.full((5,), 0.2), torch.full((7,), 1 / 7)
    with pytest.raises(ValueError, match="Explicit cluster labels"):
        SamplesLoss(backend="online")(torch.zeros(5), a, x, torch.zeros(7), b, y)
    with pytest.raises(KeyError):  # the reference dies with KeyError(None) for hausdorff (SURVEY A-13)
        SamplesLoss("hausdorff")(x, y)
    with pytest.raises(KeyError):
        SamplesLoss("nonsense")(x, y)
    with pytest.raises(KeyError):
        SamplesLoss("sinkhorn", p=3)(x, y)
    with pytest.raises(NotImplementedError):
        SamplesLoss("sinkhorn", cost=lambda u, v: u)(x, y)


def test_product_never_touches_the_oracle_or_a_cpu_fallback():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "geomloss_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "geomloss_oracle" not in text, f


# ------------------------------------------------------------------------------------------------
# geomloss.ot.solve_sample facade: argument handling (reference: sample.py:283-345, _arguments.py:14-154)
# ------------------------------------------------------------------------------------------------
def test_ot_solve_sample_argument_errors():
    from geomloss_b200 import ot

    x, y = torch.rand(5, 3), torch.rand(7, 3)
    with pytest.raises(ValueError, match="redundant"):
        ot.solve_sample(x, y, reg=0.1, blur=0.1, max_iter=3)
    with pytest.raises(ValueError, match="redundant"):
        ot.solve_sample(x, y, reg=0.1, unbalanced=1.0, reach=0.1, max_iter=3)
    with pytest.raises(ValueError, match="max_iter"):
        ot.solve_sample(x, y, reg=0.1)
    with pytest.raises(NotImplementedError):
        ot.solve_sample(x, y, reg=0.0, max_iter=3)
    with pytest.raises(ValueError):
        ot.solve_sample(x, y, reg=-1.0, max_iter=3)
    with pytest.raises(ValueError):
        ot.solve_sample(x, y, reg=0.1, unbalanced=-2.0, max_iter=3)
    with pytest.raises(NotImplementedError):
        ot.solve_sample(x, y, reg=0.1, max_iter=3, tol=1e-3)
    with pytest.raises(NotImplementedError):
        ot.solve_sample(x, y, reg=0.1, max_iter=3, method="symmetric")
    with pytest.raises(NotImplementedError):
        ot.solve_sample(x, y, reg=0.1, max_iter=3, unbalanced=1.0, unbalanced_type="TV")
    with pytest.raises(ValueError, match="same number of coordinates"):
        ot.solve_sample(x, torch.rand(7, 2), reg=0.1, max_iter=3)
    with pytest.raises(ValueError, match="X_a"):
        ot.solve_sample(torch.rand(2, 5, 3), y, reg=0.1, max_iter=3)
    with pytest.raises(ValueError, match="negative"):
        ot.solve_sample(x, y, a=-torch.ones(5), reg=0.1, max_iter=3)
    with pytest.raises(ValueError, match="shape"):
        ot.solve_sample(x, y, a=torch.ones(6), reg=0.1, max_iter=3)
    with pytest.raises(ValueError, match="do not sum up"):
        ot.solve_sample(x, y, a=torch.ones(5), b=torch.ones(7), reg=0.1, max_iter=3)
    with pytest.raises(NotImplementedError):
        ot.solve_sample_batch(x[None], y[None], reg=0.1, max_iter=3)
    # the product has no CPU path: valid arguments on CPU tensors must fail loudly, not fall back
    with pytest.raises(Exception):
        ot.solve_sample(x, y, reg=0.1, max_iter=3)


def test_ot_annealing_ladder():
    from geomloss_b200.ot import annealing_eps

    assert annealing_eps(3.0, 0.01, 1) == [0.01]
    lad = annealing_eps(3.0, 0.01, 5)
    assert len(lad) == 5 and abs(lad[0] - 3.0) < 1e-12 and abs(lad[-1] - 0.01) < 1e-12
    assert all(abs(lad[i + 1] / lad[i] - lad[1] / lad[0]) < 1e-9 for i in range(3))
    assert annealing_eps(0.5, 2.0, 3) == [2.0, 2.0, 2.0]  # reg above the squared diameter: constant ladder
    with pytest.raises(ValueError):
        annealing_eps(3.0, 0.01, 0)


# ------------------------------------------------------------------------------------------------
# ImagesBarycenter host logic on CPU: the CUDA grid softmin is replaced by the dense oracle operator (test
# infrastructure, like OracleStages in test_distributed_gloo.py); checks the iteration structure and the
# closed-form softmin_grid backward (two grid softmins) against plain autograd through the oracle.
# ------------------------------------------------------------------------------------------------
def _dense_grid_softmin(eps, p, h_a, h_b=None, h_scale_b=0.0, *, out_old=None, alpha_old=0.0, beta=1.0):
    from oracle import geomloss_oracle as O

    h = h_a if h_b is None else h_a + h_scale_b * h_b
    out = beta * O.softmin_grid_dense(eps, p, h.double()).to(h_a.dtype)
    return out if out_old is None else out + alpha_old * out_old


@pytest.mark.parametrize("backward_iterations", [0, 2])
def test_images_barycenter_host_logic(monkeypatch, backward_iterations):
    from geomloss_b200 import barycenter_images as BI
    from oracle import geomloss_oracle as O

    monkeypatch.setattr(BI, "softmin_grid", _dense_grid_softmin)
    g = torch.Generator().manual_seed(3)
    n, K = 8, 3
    imgs = torch.rand(2, K, n, n, generator=g, dtype=torch.float64) + 0.05
    imgs = imgs / imgs.sum((2, 3), keepdim=True)
    w = torch.rand(2, K, generator=g, dtype=torch.float64) + 0.2
    w = w / w.sum(1, keepdim=True)
    probe = torch.rand(2, 1, n, n, generator=g, dtype=torch.float64)
    for p in (1, 2):
        ia, wa = imgs.clone().requires_grad_(True), w.clone().requires_grad_(True)
        ib, wb = imgs.clone().requires_grad_(True), w.clone().requires_grad_(True)
        ours = BI.ImagesBarycenter(ia, wa, p=p, scaling_N=3, backward_iterations=backward_iterations)
        ref = O.images_barycenter(ib, wb, p=p, scaling_N=3, backward_iterations=backward_iterations)
        assert ours.shape == (2, 1, n, n)
        assert (ours - ref).abs().max() <= 1e-9 * ref.abs().max()
        go, gr = torch.autograd.grad((ours * probe).sum(), [wa, ia], allow_unused=True), \
            torch.autograd.grad((ref * probe).sum(), [wb, ib], allow_unused=True)
        assert (go[0] - gr[0]).abs().max() <= 1e-7 * gr[0].abs().max()
        if backward_iterations == 0:
            assert (go[1] - gr[1]).abs().max() <= 1e-7 * gr[1].abs().max()
    with pytest.raises(ValueError):
        BI.ImagesBarycenter(torch.rand(1, 2, 4, 8), torch.rand(1, 2))


# ------------------------------------------------------------------------------------------------
# ot.solve_sample host logic on CPU: ops.softmin_raw replaced by a dense fp64 stand-in with the same fused
# signature (test infrastructure); the facade must reproduce the reference's goldens.
# ------------------------------------------------------------------------------------------------
def _dense_softmin_raw(eps, x, y, h_a, h_b=None, h_scale_b=0.0, *, p=2, center=None, out_old=None, alpha_old=0.0,
                       beta=1.0, out=None, want_lse2=False):
    from oracle import geomloss_oracle as O

    h = h_a.double() if h_b is None else h_a.double() + h_scale_b * h_b.double()
    val = -eps * torch.logsumexp(h[None, :] - O.cost_matrix(x.double(), y.double(), p) / eps, dim=1) * beta
    if out_old is not None:
        val = val + alpha_old * out_old.double()
    return val.float(), None


def _dense_softmin_grad_rows(eps, x, y, h_a, h_b, h_scale_b, lse2, grad_out, *, p=2, center=None):
    """Stand-in for b200ot_softmin_bwd_x: grad_out_i * (x_i - sum_j w_ij y_j), softmax weights re-normalised by
    their own sum (lse2 only guards the kernel against overflow)."""
    from oracle import geomloss_oracle as O

    h = h_a.double() if h_b is None else h_a.double() + h_scale_b * h_b.double()
    w = torch.softmax(h[None, :] - O.cost_matrix(x.double(), y.double(), p) / eps, dim=1)
    return (grad_out.double()[:, None] * (x.double() - w @ y.double())).float()


@pytest.mark.parametrize("idx", range(8))
def test_ot_solve_sample_gradients_host_logic(monkeypatch, idx):
    """d value / d (X_a, X_b, a, b) of the facade (last update differentiated w.r.t. BOTH clouds: row-gradient
    kernel + the same kernel on the swapped problem for the columns) against the real reference's autograd."""
    from conftest import load_golden
    from geomloss_b200 import ops, ot

    monkeypatch.setattr(ops, "softmin_raw", _dense_softmin_raw)
    monkeypatch.setattr(ops, "softmin_grad_rows", _dense_softmin_grad_rows)
    z = load_golden(f"ot_sample_case{idx:02d}")
    kw = {k[3:]: float(z[k]) for k in z if k.startswith("kw_")}
    kw["max_iter"] = int(kw["max_iter"])
    if "debias" in kw:
        kw["debias"] = bool(kw["debias"])
    n, m = z["X_a"].shape[0], z["X_b"].shape[0]
    a0 = torch.from_numpy(z["a"]) if "a" in z else torch.full((n,), 1.0 / n)
    b0 = torch.from_numpy(z["b"]) if "b" in z else torch.full((m,), 1.0 / m)
    leaves = [t.clone().requires_grad_(True) for t in (torch.from_numpy(z["X_a"]), torch.from_numpy(z["X_b"]), a0, b0)]
    res = ot.solve_sample(leaves[0], leaves[1], a=leaves[2], b=leaves[3], **kw)
    assert abs(res.value.item() - float(z["value_f64"])) <= 2e-5 * abs(float(z["value_f64"]))
    grads = torch.autograd.grad(res.value, leaves)
    for g, name in zip(grads, ("grad_X_a", "grad_X_b", "grad_a", "grad_b")):
        ref = z[name + "_f64"]
        np.testing.assert_allclose(g.numpy(), ref, atol=2e-4 * float(np.abs(ref).max()), err_msg=name)
    assert not res.potential_a.requires_grad and not res.marginal_a.requires_grad  # accessors are detached
    # without requires_grad (or under no_grad) nothing is attached
    with torch.no_grad():
        assert not ot.solve_sample(leaves[0], leaves[1], a=leaves[2], b=leaves[3], **kw).value.requires_grad


@pytest.mark.parametrize("idx", range(8))
def test_ot_solve_sample_host_logic(monkeypatch, idx):
    from conftest import load_golden
    from geomloss_b200 import ops, ot

    monkeypatch.setattr(ops, "softmin_raw", _dense_softmin_raw)
    z = load_golden(f"ot_sample_case{idx:02d}")
    kw = {k[3:]: float(z[k]) for k in z if k.startswith("kw_")}
    kw["max_iter"] = int(kw["max_iter"])
    if "debias" in kw:
        kw["debias"] = bool(kw["debias"])
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    res = ot.solve_sample(t("X_a"), t("X_b"), a=t("a") if "a" in z else None, b=t("b") if "b" in z else None, **kw)
    ref = float(z["value_f64"])
    assert abs(res.value.item() - ref) <= 2e-5 * abs(ref)
    names = ["potential_a", "potential_b", "marginal_a", "marginal_b", "plan"] + (
        ["potential_aa", "potential_bb"] if kw.get("debias") else [])
    for name in names:
        r = z[name + "_f64"]
        tol = 1e-5 * max(1.0, float(np.abs(r).max())) if "potential" in name else 1e-4 * float(np.abs(r).max())
        np.testing.assert_allclose(getattr(res, name).numpy(), r, atol=tol, err_msg=name)
    # the plan as an operator (two softmins per column, signed right-hand sides), against the dense plan
    P = torch.from_numpy(z["plan_f64"])
    s = torch.randn(P.shape[1], 3, generator=torch.Generator().manual_seed(idx), dtype=torch.float64)
    out = (res.plan_operator @ s.float()).double()
    assert (out - P @ s).abs().max() <= 1e-4 * (P @ s.abs()).max()
    res.cache_clear()
    assert abs(res.value.item() - ref) <= 2e-5 * abs(ref)


# ------------------------------------------------------------------------------------------------
# numerics of the tensor-core operand format (csrc/tcconv.cuh): X = h + l in fp16, products hh + hl + lh
# ------------------------------------------------------------------------------------------------
def test_fp16_two_term_split_error_bound():
    """CPU emulation of the split the tensor-core kernels use for 8 < D <= 64: the exponent
    S = X.Y - |X|^2/2 - |Y|^2/2 formed from two fp16 terms per coordinate and three cross products (exact
    accumulation emulated in fp64) stays within 2^-20 |X||Y| of the fp64 value — i.e. the fp32 accumulator, not the
    16-bit operands, bounds the accuracy of the kernel value."""
    g = torch.Generator().manual_seed(0)
    D, n = 64, 300
    x = torch.rand(n, D, generator=g, dtype=torch.float64)
    y = torch.rand(n, D, generator=g, dtype=torch.float64)
    for blur in (2.0, 0.7, 0.3):
        s = np.sqrt(np.log2(np.e)) / blur
        X, Y = ((x - 0.5) * s).float(), ((y - 0.5) * s).float()
        Xh, Yh = X.half(), Y.half()
        Xl, Yl = (X - Xh.float()).half(), (Y - Yh.float()).half()
        dot = Xh.double() @ Yh.double().T + Xh.double() @ Yl.double().T + Xl.double() @ Yh.double().T
        exact = X.double() @ Y.double().T
        bound = 2.0**-20 * X.double().norm(dim=1)[:, None] * Y.double().norm(dim=1)[None, :]
        err = (dot - exact).abs()
        assert bool((err <= bound).all()), (blur, float((err / bound).max()))
        # relative error of the kernel value 2^S: below 1e-5 down to blur = .3 (|X|^2 ~ 86)
        assert float(err.max()) * np.log(2) < 1e-5, (blur, float(err.max()))


# ------------------------------------------------------------------------------------------------
# sinkhorn_images.sinkhorn_divergence host logic on CPU (pyramid, jumps, up-sampling, value formulas) with the
# dense grid operator standing in for b200ot_softmin_grid
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 1, 16, 16), (1, 2, 8, 8, 8)])
@pytest.mark.parametrize("kw", [dict(p=2), dict(p=1, reach=0.3), dict(p=2, debias=False, scaling=0.7)])
def test_sinkhorn_images_host_logic(monkeypatch, shape, kw):
    from geomloss_b200 import sinkhorn_images as SI
    from oracle import geomloss_oracle as O

    monkeypatch.setattr(SI, "softmin_grid", _dense_grid_softmin)
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.rand(*shape, generator=g, dtype=torch.float64) + 0.01
    b = torch.rand(*shape, generator=g, dtype=torch.float64) + 0.01
    dims = tuple(range(2, len(shape)))
    a, b = a / a.sum(dims, keepdim=True), 1.2 * b / b.sum(dims, keepdim=True) if "reach" in kw else b / b.sum(dims, keepdim=True)
    ours = SI.sinkhorn_divergence(a, b, **kw)
    ref = O.sinkhorn_images(a, b, **kw)
    assert ours.shape == (shape[0],)
    assert (ours - ref).abs().max() <= 1e-9 * ref.abs().max() + 1e-15
    F, G = SI.sinkhorn_divergence(a, b, potentials=True, **kw)
    Fr, Gr = O.sinkhorn_images(a, b, potentials=True, **kw)
    assert F.shape == a.shape and (F - Fr).abs().max() <= 1e-9 and (G - Gr).abs().max() <= 1e-9
    with pytest.raises(ValueError):
        SI.sinkhorn_divergence(a, b, scaling=0.3)


# ---------------------------------------------------------------------------------------------------------
# ranges mode descriptors (host logic of geomloss_b200/ranges.py; the C library is only asked for tile shapes)
# ---------------------------------------------------------------------------------------------------------
def _pair_mask(prob, n_rows, n_cols):
    src = prob.layout.src.tolist()
    pieces = prob.pieces.tolist()
    mask = torch.zeros(n_rows, n_cols, dtype=torch.int32)
    for r0, nr, p0, p1 in prob.seg.tolist():
        for c0, nc in pieces[p0:p1]:
            cols = [j for j in src[c0:c0 + nc] if j >= 0]
            mask[r0:r0 + nr, cols] += 1
    return mask


@pytest.mark.parametrize("variant", [0, 1])
def test_ranges_descriptors_list_exactly_the_kept_cluster_pairs(variant):
    """segments x pieces == keep[lab_rows][:, lab_cols], each pair exactly once; pieces are aligned and tile-sized;
    a column-sharded build partitions the pairs."""
    from geomloss_b200 import ranges

    g = torch.Generator().manual_seed(variant)
    R, C = 23, 17
    row_counts = torch.randint(1, 700, (R,), generator=g)
    col_counts = torch.randint(1, 1500, (C,), generator=g)
    keep = torch.rand(R, C, generator=g) < 0.4
    keep[3] = False  # a row cluster that keeps nothing
    keep[5] = True   # ... and one that keeps everything (one long run, cut into tile-sized pieces)
    lay = ranges.ColumnLayout(col_counts)
    n_rows, n_cols = int(row_counts.sum()), int(col_counts.sum())
    lab_r = torch.repeat_interleave(torch.arange(R), row_counts)
    lab_c = torch.repeat_interleave(torch.arange(C), col_counts)
    want = keep[lab_r][:, lab_c].to(torch.int32)
    max_rows, max_cols, align = ranges.shape(variant)
    prob = ranges.build_problem(keep, row_counts, lay, variant=variant)
    assert torch.equal(_pair_mask(prob, n_rows, n_cols), want)
    assert prob.seg.dtype == torch.int32 and prob.pieces.dtype == torch.int32
    assert int(prob.seg[:, 1].max()) <= max_rows and int(prob.seg[:, 1].min()) >= 1
    assert int(prob.pieces[:, 1].max()) <= max_cols
    assert bool((prob.pieces % align == 0).all())
    assert abs(prob.density - want.double().mean().item()) < 1e-12
    # padding slots are neutral, real columns appear exactly once
    src = lay.src
    assert sorted(src[src >= 0].tolist()) == list(range(n_cols)) and lay.n_slots % align == 0
    # column-sharded: the ranks' problems partition the kept pairs and are roughly balanced
    world = 3
    parts = [ranges.build_problem(keep, row_counts, lay, variant=variant, rank=r, world=world) for r in range(world)]
    masks = [_pair_mask(p, n_rows, n_cols) for p in parts]
    assert torch.equal(sum(masks), want)
    loads = [float(m.sum()) for m in masks]
    assert max(loads) <= 0.6 * sum(loads)


def test_batch_problem_is_block_diagonal():
    from geomloss_b200 import ranges

    B, N, M = 3, 150, 70
    prob = ranges.batch_problem(B, N, M, "cpu")
    want = torch.block_diag(*[torch.ones(N, M, dtype=torch.int32)] * B)
    assert torch.equal(_pair_mask(prob, B * N, B * M), want)
    assert prob.density == pytest.approx(1.0 / B)


def test_uniform_weights_are_cached_per_shape():
    """generate_weights (samples_loss.py:328-335 in the reference): 1/N per point; the engine keeps one read-only tensor
    per (shape, device, dtype) instead of rebuilding it on every call."""
    from geomloss_b200 import SamplesLoss

    L = SamplesLoss("sinkhorn")
    x = torch.zeros(7, 3)
    w = L.generate_weights(x)
    assert w.shape == (7,) and torch.allclose(w, torch.full((7,), 1 / 7)) and not w.requires_grad
    assert L.generate_weights(torch.ones(7, 2)) is w  # same leading shape, device, dtype
    wb = L.generate_weights(torch.zeros(4, 5, 3))
    assert wb.shape == (4, 5) and torch.allclose(wb, torch.full((4, 5), 0.2))
    assert L.generate_weights(torch.zeros(7, 3, dtype=torch.float64)).dtype == torch.float64
    with pytest.raises(ValueError):
        L.generate_weights(torch.zeros(3))
