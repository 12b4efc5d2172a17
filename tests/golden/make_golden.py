"""Generate the golden fixtures in this directory by running the REAL reference.

Run once in the build container (the reference is mounted read-only at /root/reference and cannot travel
to the GPU box):

    python tests/golden/make_golden.py

It imports ``geomloss`` from /root/reference/src (commit 00e493f, v0.3.1), evaluates
``SamplesLoss(..., backend="tensorized")`` and the operator-level functions of the hot path on seeded
inputs, and stores inputs + outputs as small ``.npz`` files.  The tests compare (a) the CPU oracle and
(b) the CUDA engine against these files.  Nothing here is imported by the product.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("GEOMLOSS_REFERENCE", "/root/reference/src")
sys.path.insert(0, REF)
import geomloss  # noqa: E402
from geomloss import SamplesLoss  # noqa: E402
from geomloss._legacy import sinkhorn_divergence as ref_sd  # noqa: E402
from geomloss._legacy import sinkhorn_samples as ref_ss  # noqa: E402
from geomloss._legacy import kernel_samples as ref_ks  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def np32(t):
    return t.detach().cpu().numpy()


def clouds(seed, n, m, d, dtype=torch.float32, weights="uniform", batch=None):
    g = torch.Generator().manual_seed(seed)
    shape_x = (n, d) if batch is None else (batch, n, d)
    shape_y = (m, d) if batch is None else (batch, m, d)
    x = torch.rand(*shape_x, generator=g, dtype=torch.float64)
    y = torch.rand(*shape_y, generator=g, dtype=torch.float64) * 0.9 + 0.15
    if weights == "uniform":
        a = torch.ones(shape_x[:-1], dtype=torch.float64) / n
        b = torch.ones(shape_y[:-1], dtype=torch.float64) / m
    else:
        a = torch.rand(shape_x[:-1], generator=g, dtype=torch.float64) + 0.1
        b = torch.rand(shape_y[:-1], generator=g, dtype=torch.float64) + 0.1
        if weights == "random_with_zero":
            a[..., 0] = 0.0
            b[..., -1] = 0.0
        a = a / a.sum(-1, keepdim=True)
        b = b / b.sum(-1, keepdim=True)
        if weights == "unbalanced":
            b = b * 1.3
    return a.to(dtype), x.to(dtype), b.to(dtype), y.to(dtype)


def run_loss(a, x, b, y, grads=True, **kw):
    """value, potentials and gradients of the reference's tensorized SamplesLoss."""
    out = {}
    L = SamplesLoss(backend="tensorized", **kw)
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    val = L(ag, xg, bg, yg)
    out["value"] = np32(val)
    if grads:
        ga, gx, gb, gy = torch.autograd.grad(val.sum(), [ag, xg, bg, yg])
        out.update(grad_a=np32(ga), grad_x=np32(gx), grad_b=np32(gb), grad_y=np32(gy))
    Lp = SamplesLoss(backend="tensorized", potentials=True, **kw)
    F, G = Lp(a, x, b, y)
    out.update(pot_f=np32(F), pot_g=np32(G))
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    meta = dict(reference_version=geomloss.__version__, torch=torch.__version__)
    print(meta)

    # ---- cfg 1 of BASELINE.json: N=M=1000, D=3, blur=.05, seed 0 --------
    torch.manual_seed(0)
    x = torch.rand(1000, 3)
    y = torch.rand(1000, 3)
    a = torch.ones(1000) / 1000
    b = torch.ones(1000) / 1000
    r32 = run_loss(a, x, b, y, loss="sinkhorn", p=2, blur=0.05)
    r64 = run_loss(a.double(), x.double(), b.double(), y.double(), loss="sinkhorn", p=2, blur=0.05)
    save("cfg1_sinkhorn_n1000", x=np32(x), y=np32(y), a=np32(a), b=np32(b),
         **{k + "_f32": v for k, v in r32.items()}, **{k + "_f64": v for k, v in r64.items()})
    print("cfg1 value", r32["value"], r64["value"])

    # ---- Sinkhorn matrix of small cases ------------------------------------------------------
    cases = []
    idx = 0
    for (n, m, d) in [(37, 53, 3), (64, 64, 2), (50, 41, 1), (45, 60, 5)]:
        for p in (1, 2):
            for reach in (None, 0.3):
                for debias in (True, False):
                    wts = "unbalanced" if reach is not None else ("random_with_zero" if idx % 2 else "random")
                    a_, x_, b_, y_ = clouds(100 + idx, n, m, d, weights=wts)
                    kw = dict(loss="sinkhorn", p=p, blur=0.1 if p == 2 else 0.05, reach=reach, debias=debias,
                              scaling=0.6)
                    r = run_loss(a_, x_, b_, y_, **kw)
                    # the same reference code on the same (fp32-representable) inputs, evaluated in fp64:
                    # the fp32 run has visible artefacts for p=1 (noisy sqrt(clamp) on near-zero distances)
                    r64 = run_loss(a_.double(), x_.double(), b_.double(), y_.double(), **kw)
                    cases.append((kw, n, m, d))
                    arrays = dict(a=np32(a_), x=np32(x_), b=np32(b_), y=np32(y_),
                                  p=np.int64(p), blur=np.float64(kw["blur"]),
                                  reach=np.float64(-1 if reach is None else reach), debias=np.int64(debias),
                                  scaling=np.float64(0.6), **r, **{k + "_f64": v for k, v in r64.items()})
                    save(f"sinkhorn_case{idx:02d}", **arrays)
                    idx += 1

    # ---- batched input (B=2) -----------------------------------------------------------------
    a_, x_, b_, y_ = clouds(7, 40, 30, 3, weights="random", batch=2)
    r = run_loss(a_, x_, b_, y_, loss="sinkhorn", p=2, blur=0.1)
    r64 = run_loss(a_.double(), x_.double(), b_.double(), y_.double(), loss="sinkhorn", p=2, blur=0.1)
    save("sinkhorn_batched", a=np32(a_), x=np32(x_), b=np32(b_), y=np32(y_), **r,
         **{k + "_f64": v for k, v in r64.items()})

    # ---- kernel MMDs ----------------------------------------------------------------------------
    for k, name in enumerate(["gaussian", "laplacian", "energy"]):
        for j, (n, m, d, blur) in enumerate([(48, 61, 3, 0.3), (33, 20, 5, 0.7)]):
            a_, x_, b_, y_ = clouds(500 + 10 * k + j, n, m, d, weights="random")
            r = run_loss(a_, x_, b_, y_, loss=name, blur=blur)
            r64 = run_loss(a_.double(), x_.double(), b_.double(), y_.double(), loss=name, blur=blur)
            save(f"kernel_{name}_{j}", a=np32(a_), x=np32(x_), b=np32(b_), y=np32(y_), blur=np.float64(blur),
                 **r, **{k + "_f64": v for k, v in r64.items()})

    # ---- operator level: softmin_tensorized on arbitrary h -------------------------------------
    a_, x_, b_, y_ = clouds(900, 70, 90, 3, weights="random")
    g = torch.Generator().manual_seed(901)
    pot = (torch.rand(90, generator=g) - 0.5) * 0.2
    ops = {}
    for p in (1, 2):
        C = ref_ss.cost_routines[p](x_.unsqueeze(0), y_.unsqueeze(0))
        C64 = ref_ss.cost_routines[p](x_.double().unsqueeze(0), y_.double().unsqueeze(0))
        for e, eps in enumerate([1.0, 0.05, 0.003]):
            h = ref_sd.log_weights(b_) + pot / eps
            ops[f"softmin_p{p}_eps{e}"] = np32(ref_ss.softmin_tensorized(eps, C, h.unsqueeze(0)))[0]
            h64 = ref_sd.log_weights(b_.double()) + pot.double() / eps
            ops[f"softmin_p{p}_eps{e}_f64"] = np32(ref_ss.softmin_tensorized(eps, C64, h64.unsqueeze(0)))[0]
    save("softmin_operator", x=np32(x_), y=np32(y_), b=np32(b_), pot=np32(pot),
         eps=np.array([1.0, 0.05, 0.003]), **ops)

    # ---- schedule ---------------------------------------------------------------------------------
    sched = {}
    for i, (p, diam, blur, scaling) in enumerate([(2, 1.7306, 0.05, 0.5), (2, 3.0 ** 0.5, 0.01, 0.9),
                                                  (1, 2.5, 0.05, 0.7), (2, 1.0, 0.01, 0.5)]):
        sched[f"args{i}"] = np.array([p, diam, blur, scaling], dtype=np.float64)
        sched[f"eps{i}"] = np.array(ref_sd.epsilon_schedule(p, diam, blur, scaling), dtype=np.float64)
    save("epsilon_schedule", **sched)

    # ---- kernel matrices (operator level) ---------------------------------------------------------
    km = {}
    for name in ("gaussian", "laplacian", "energy"):
        km[name] = np32(ref_ks.kernel_routines[name](x_, y_, blur=0.25))
    save("kernel_operator", x=np32(x_), y=np32(y_), blur=np.float64(0.25), **km)


if __name__ == "__main__":
    main()
