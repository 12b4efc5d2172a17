"""Where does the HOST time of a small SamplesLoss call go?  cProfile over repeated forward+backward calls at N = 1000
(the regime is host-bound: the device work of one call is ~0.2 ms).  python tools/profile_small_host.py [N] [loss]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from geomloss_b200 import SamplesLoss  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
loss = sys.argv[2] if len(sys.argv) > 2 else "sinkhorn"
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 3, generator=g).to(dev).requires_grad_(True)
y = torch.rand(N, 3, generator=g).to(dev)
L = SamplesLoss(loss, p=2, blur=0.05)


def call():
    v = L(x, y)
    (gx,) = torch.autograd.grad(v, x)
    return v


for _ in range(20):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    call()
torch.cuda.synchronize()
print(f"{loss} N={N}: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms per fwd+bwd")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    call()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
