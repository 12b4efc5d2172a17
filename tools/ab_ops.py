"""A/B timing of the C-ABI operators between two builds of the library.

    B200OT_LIB=build/libb200ot_A.so python tools/ab_ops.py --tag A > a.jsonl
    python tools/ab_ops.py --tag B > b.jsonl

Every operator is timed alone with CUDA events (best of --reps after one warm-up) on N = M = --n points, through the
public ops layer — the same calls the host loop makes.  Used to pick compile-time tile parameters (softmin chunk
length, rowsum unroll) for all (D, p) at once instead of the D = 3 study of tools/explore.cu.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomloss_b200 import _lib, ops  # noqa: E402


def best_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=400000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N = M = args.n
    g = torch.Generator().manual_seed(0)
    for D in (1, 2, 3, 4, 6, 8):
        x = torch.rand(N, D, generator=g).to(dev)
        y = torch.rand(M, D, generator=g).to(dev)
        h = (torch.rand(M, generator=g) * 0.1).to(dev)
        w = (torch.rand(M, generator=g) / M).to(dev)
        go = torch.rand(N, generator=g).to(dev)
        rows = []
        for p in (2, 1):
            eps = 0.01 ** p
            _, lse2 = ops.softmin_raw(eps, x, y, h, p=p, want_lse2=True)
            rows.append((f"softmin_fwd p={p}", lambda p=p, eps=eps: ops.softmin_raw(eps, x, y, h, p=p)))
            rows.append((f"softmin_bwd p={p}", lambda p=p, eps=eps, lse2=lse2: ops.softmin_grad_rows(
                eps, x, y, h, None, 0.0, lse2, go, p=p)))
        for kind in ("gaussian", "laplacian", "energy"):
            rows.append((f"conv_fwd {kind}", lambda kind=kind: ops.kernel_conv_raw(kind, x, y, w, 0.05)))
            rows.append((f"conv_bwd {kind}", lambda kind=kind: ops.kernel_conv_grad_rows(kind, x, y, w, 0.05, go)))
        for name, fn in rows:
            ms = best_ms(fn, args.reps)
            print(json.dumps(dict(tag=args.tag, lib=os.path.basename(_lib.LIB_PATH), op=name, D=D, N=N, M=M,
                                  ms=round(ms, 3), Tpairs_s=round(N * M / ms * 1e-9, 3))), flush=True)


if __name__ == "__main__":
    main()
