"""A/B timing of the small-problem kernels between builds of the library ($B200OT_LIB): device time of one whole
Sinkhorn descent (b200ot_sinkhorn_loop_small: n_eps + 1 launches) and of the fused kernel-MMD forward, CUDA events,
best of 20.   python tools/ab_small.py --tag W16T512"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomloss_b200 import _lib, kernel_small, ops, sinkhorn_small  # noqa: E402
from geomloss_b200.sinkhorn import scaling_parameters  # noqa: E402


def best_us(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    for B, N in ((1, 200), (1, 1000), (1, 3000), (1, 6000), (64, 500), (256, 100)):
        x = torch.rand(B, N, 3, generator=g).to(dev)
        y = torch.rand(B, N, 3, generator=g).to(dev)
        a = torch.full((B, N), 1.0 / N, device=dev)
        _, _, eps_list, rho = scaling_parameters(x, y, 2, 0.05, None, 1.7320508, 0.5)
        t = best_us(lambda: sinkhorn_small._descent(x, y, a, a, eps_list, rho, 2, True))
        kid = ops.KERNEL_KINDS["gaussian"]
        t2 = best_us(lambda: kernel_small._forward(kid, a, x, a, y, 0.1, False))
        print(json.dumps(dict(tag=args.tag, lib=os.path.basename(_lib.LIB_PATH), B=B, N=N, n_launch=len(eps_list) + 1,
                              descent_us=round(t, 1), us_per_iteration=round(t / (len(eps_list) + 1), 2),
                              mmd_fwd_us=round(t2, 1))), flush=True)


if __name__ == "__main__":
    main()
