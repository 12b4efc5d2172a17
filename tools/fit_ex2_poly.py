"""Fit the degree-5 polynomial used by csrc/common.cuh::ex2_poly (2^f on [-0.5, 0.5]).

Iteratively re-weighted least squares on the relative error (a poor man's Remez) with the
constant term pinned to 1 so that 2^0 == 1 exactly.  Prints the coefficients low -> high order
and the achieved max relative error in exact and in fp32 Horner arithmetic.
"""
import numpy as np


def fit(deg: int):
    f = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5
    y = 2.0**f
    w = np.ones_like(f)
    for _ in range(60):
        A = np.vander(f, deg + 1, increasing=True)
        coef = np.linalg.lstsq(A[:, 1:] * (w / y)[:, None], (y - 1) * w / y, rcond=None)[0]
        coef = np.r_[1.0, coef]
        err = np.abs(A @ coef - y) / y
        w = w * (1 + 3 * err / err.max())
        w /= w.mean()
    return coef, err.max()


if __name__ == "__main__":
    for d in (4, 5, 6):
        c, e = fit(d)
        f = np.linspace(-0.5, 0.5, 200001).astype(np.float32)
        p = np.float32(c[-1]) * np.ones_like(f)
        for k in range(d - 1, -1, -1):
            p = (p * f + np.float32(c[k])).astype(np.float32)
        ref = 2.0 ** f.astype(np.float64)
        e32 = np.max(np.abs(p.astype(np.float64) - ref) / ref)
        print(f"deg {d}: max rel err {e:.3e} (exact) {e32:.3e} (fp32 Horner)")
        print("   ", ", ".join(f"{v:.10e}" for v in c))
