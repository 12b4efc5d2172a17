"""One forward call of BASELINE configs[3] at N = M = 1e6 (blur .01, scaling .5, truncate 5) — the workload
tools/gpu_round2_q.sh puts under `ncu --metrics gpu__time_duration.sum` to see where its 1.4 s go."""
import sys
import time

import torch

sys.path.insert(0, ".")
from geomloss_b200 import SamplesLoss  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 3, generator=g).cuda()
y = torch.rand(N, 3, generator=g).cuda()
L = SamplesLoss("sinkhorn", p=2, blur=0.01, scaling=0.5, truncate=5, backend="multiscale")
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v = L(x, y).item()
    print(f"call {rep}: {time.perf_counter() - t0:.3f} s, value {v:.6e}", flush=True)
