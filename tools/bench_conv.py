"""Time the kernel-convolution entry points (CUDA events, L2 flushed between reps).
    python tools/bench_conv.py [N] [D ...]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from geomloss_b200 import ops  # noqa: E402  (bench.py, imported below for its NVML clock sampler, lives at the repo root)

dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dims = [int(a) for a in sys.argv[2:]] or [3, 16, 32, 64]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for D in dims:
    g = torch.Generator().manual_seed(D)
    x = torch.rand(N, D, generator=g).to(dev)
    y = torch.rand(N, D, generator=g).to(dev)
    w = torch.full((N,), 1.0 / N, device=dev)
    c = ops.default_center(x, y)
    for kind in (("gaussian",) if D > 8 else ("gaussian", "laplacian", "energy")):
        blur = 2.0 if D > 8 else 0.1
        ops.kernel_conv_raw(kind, x, y, w, blur, center=c)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.kernel_conv_raw(kind, x, y, w, blur, center=c)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        # sustained run with the SM clock sampled under load (tensor-heavy kernels do not hold 1965 MHz: a
        # pairs/s figure only compares to a per-clock roofline at the clock it was measured at)
        from bench import ClockSampler  # noqa: E402

        reps = max(4, int(1500 / max(ms, 1e-3)))
        with ClockSampler(0, period=0.02) as cs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = ops.kernel_conv_raw(kind, x, y, w, blur, center=c)
            e1.record()
            torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        clk = cs.summary()
        print(json.dumps({"conv": kind, "N": N, "M": N, "D": D, "ms": round(ms, 3),
                          "Tpairs_s": round(N * N / (ms * 1e-3) / 1e12, 3), "sustained_ms": round(sus, 3),
                          "sustained_Tpairs_s": round(N * N / (sus * 1e-3) / 1e12, 3), "sm_mhz": clk.get("sm_mhz"),
                          "sm_min_mhz": clk.get("sm_min_mhz"), "power_w_max": clk.get("power_w_max"),
                          "reasons": clk.get("reasons"), "sum": float(out.sum())}))

# forward + backward of the gaussian MMD at D = 64 (BASELINE configs[2] protocol: L = Loss(x, y); L.backward())
if 64 in dims:
    import time

    from geomloss_b200 import SamplesLoss

    n = N  # (round 1 capped this at 4e5; BASELINE configs[2] is N = M = 1e6)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 64, generator=g).to(dev).requires_grad_(True)
    y = torch.rand(n, 64, generator=g).to(dev)
    L = SamplesLoss("gaussian", blur=2.0)
    for phase in ("fwd", "fwd+bwd"):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            val = L(x, y)
            if phase == "fwd+bwd":
                (gx,) = torch.autograd.grad(val, x)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        # forward: 3 matvecs (xx, yy, xy); backward adds d/dx of the xx and xy terms (2 row-gradient reductions)
        print(json.dumps({"loss": "gaussian MMD", "N": n, "D": 64, "phase": phase, "s": round(dt, 4),
                          "pairs": (3 if phase == "fwd" else 5) * n * n,
                          "Tpairs_s": round((3 if phase == "fwd" else 5) * n * n / dt / 1e12, 3)}))
